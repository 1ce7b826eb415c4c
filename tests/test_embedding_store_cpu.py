"""Embedding store: files and contents must be what the reference loop writes
(chexpert-get-embedding.py:68-113, batch size 1, checkpoint every 5000) and what glue_dataset.py / Trainer read."""
import os

import pytest
import torch
from torch.utils.data import ConcatDataset, DataLoader, TensorDataset

from incremental_multimodal_medical_learning_ii_b200 import embedding_store as ES
from incremental_multimodal_medical_learning_ii_b200.extraction import save_embedding_chunks


def _reference_loop(emb, lab, out_dir, interval):
    """The reference's accumulation / checkpoint logic restated (chexpert-get-embedding.py:68-113), batch size 1."""
    os.makedirs(out_dir, exist_ok=True)
    e_list, l_list = torch.empty(0, 128), torch.empty(0, 5)
    files = []
    counter = 0
    for i in range(emb.shape[0]):
        e_list = torch.cat([e_list, emb[i:i + 1]], dim=0)
        l_list = torch.cat([l_list, lab[i:i + 1]], dim=0)
        counter += 1
        if counter % interval == 0:
            path = os.path.join(out_dir, f"embeddings_dataset_{counter}.pt")
            torch.save(TensorDataset(e_list, l_list), path)
            files.append(path)
            e_list, l_list = torch.empty(0, 128), torch.empty(0, 5)
    if len(e_list) > 0:
        path = os.path.join(out_dir, "embeddings_dataset_final.pt")
        torch.save(TensorDataset(e_list, l_list), path)
        files.append(path)
    return files


@pytest.mark.parametrize("n,interval,blocks", [(23, 5, (7, 1, 9, 6)), (20, 5, (20,)), (3, 5, (1, 1, 1)), (12, 4, (5, 7))])
def test_async_writer_matches_reference_loop(tmp_path, n, interval, blocks):
    g = torch.Generator().manual_seed(n)
    emb = torch.randn(n, 128, generator=g)
    lab = (torch.rand(n, 5, generator=g) > 0.5).float()
    ref_files = _reference_loop(emb, lab, str(tmp_path / "ref"), interval)
    with ES.AsyncChunkWriter(str(tmp_path / "ours"), chunk=interval) as w:
        pos = 0
        for b in blocks:
            w.add(emb[pos:pos + b], lab[pos:pos + b])
            pos += b
        assert pos == n
    assert [os.path.basename(p) for p in w.paths] == [os.path.basename(p) for p in ref_files]
    for a, b in zip(w.paths, ref_files):
        da, db = torch.load(a, weights_only=False), torch.load(b, weights_only=False)
        assert torch.equal(da.tensors[0], db.tensors[0]) and torch.equal(da.tensors[1], db.tensors[1])
    # the synchronous writer of extraction.py produces the same files
    sync = save_embedding_chunks(emb, lab, str(tmp_path / "sync"), chunk=interval)
    assert [os.path.basename(p) for p in sync] == [os.path.basename(p) for p in ref_files]


def test_chunk_order_is_numeric_and_glue_is_trainer_compatible(tmp_path):
    n, interval = 57, 5                       # 11 numbered chunks: "10", "15" ... sort numerically, not as strings
    emb = torch.arange(n, dtype=torch.float32).unsqueeze(1).repeat(1, 128)
    lab = torch.zeros(n, 5)
    with ES.AsyncChunkWriter(str(tmp_path), chunk=interval) as w:
        w.add(emb, lab)
    ds = ES.load_embedding_chunks(str(tmp_path))
    assert isinstance(ds, ConcatDataset) and len(ds) == n
    flat = ES.concat_to_tensors(ds)           # Trainer.concat_to_tensor_dataloader's walk over .datasets[i].tensors
    assert torch.equal(flat.tensors[0][:, 0], torch.arange(n, dtype=torch.float32))
    # glue_dataset.py itself only globs the numbered chunks
    assert len(ES.load_embedding_chunks(str(tmp_path), include_final=False)) == n - n % interval
    glued = ES.glue_embedding_chunks(str(tmp_path))
    loaded = torch.load(glued, weights_only=False)   # Trainer.py:221-235
    xb, yb = next(iter(DataLoader(loaded, batch_size=16, shuffle=False)))
    assert xb.shape == (16, 128) and yb.shape == (16, 5) and torch.equal(xb[:, 0], torch.arange(16.0))


def test_writer_validates_shapes(tmp_path):
    w = ES.AsyncChunkWriter(str(tmp_path), chunk=4)
    with pytest.raises(ValueError):
        w.add(torch.zeros(2, 64), torch.zeros(2, 5))
    with pytest.raises(ValueError):
        w.add(torch.zeros(2, 128), torch.zeros(3, 5))
    assert w.close() == []


def test_load_missing_store(tmp_path):
    with pytest.raises(FileNotFoundError):
        ES.load_embedding_chunks(str(tmp_path))
