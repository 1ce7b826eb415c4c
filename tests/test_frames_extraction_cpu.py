"""Synthetic frame generator + sharded extraction driver (world_size-2 gloo on CPU, no GPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from incremental_multimodal_medical_learning_ii_b200 import frames as FR
from incremental_multimodal_medical_learning_ii_b200.extraction import (extract_shard, gather_shards,
                                                                          save_embedding_chunks, shard_range)


def test_frames_are_index_addressed_and_deterministic():
    for kind in ("iid", "structured"):
        a = FR.synthetic_frames_u8(0, 6, 64, kind=kind, seed=0)
        assert a.dtype == torch.uint8 and a.shape == (6, 1, 64, 64)
        b = torch.cat([FR.synthetic_frames_u8(0, 2, 64, kind=kind, seed=0),
                       FR.synthetic_frames_u8(2, 4, 64, kind=kind, seed=0)])
        assert torch.equal(a, b)                        # any batch split / shard yields the same frames
        assert not torch.equal(a[0], a[1])
        assert not torch.equal(a, FR.synthetic_frames_u8(0, 6, 64, kind=kind, seed=1))
    x = FR.frames_as_reference_input(a)
    assert x.shape == (6, 3, 64, 64) and x.dtype == torch.float32 and 0 <= x.min() and x.max() <= 1
    assert torch.equal(x[:, 0], x[:, 2])
    with pytest.raises(ValueError):
        FR.synthetic_frames_u8(0, 1, 64, kind="nope")


def test_shard_ranges_tile_the_dataset():
    for n in (0, 1, 7, 224000, 224001, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(224000, 7, 8) == (196000, 224000)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _stub_embed(frames):
    """Cheap deterministic stand-in for the model: per-frame statistics, so order mistakes are visible."""
    f = frames.float().flatten(1)
    g = torch.stack([f.mean(1), f.std(1), f[:, 0], f[:, -1]], dim=1)
    return {"global": g, "prob": torch.sigmoid(g[:, :3] / 255), "pred": (g[:, :3] > 100).to(torch.uint8)}


def _source(first, count):
    return FR.synthetic_frames_u8(first, count, 32, kind="iid", seed=5)


def _worker(rank, world, n, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = extract_shard(_stub_embed, _source, n, batch_size=4, rank=rank, world_size=world)
        calls = []
        real = dist.all_gather_into_tensor
        dist.all_gather_into_tensor = lambda out, inp, *a, **k: (calls.append(tuple(inp.shape)), real(out, inp, *a, **k))[1]
        full = gather_shards(local, n, rank, world)
        dist.all_gather_into_tensor = real
        assert len(calls) == 1, calls            # ONE collective for embeddings + probabilities + labels
        torch.save((rank, full, local["range"].tolist()), os.path.join(outdir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [23, 16, 1])        # n = 1: rank 1 owns an EMPTY shard and must still enter the collective
def test_two_rank_extraction_matches_sequential(n, tmp_path):
    world = 2
    seq = extract_shard(_stub_embed, _source, n, batch_size=5)
    ctx = mp.get_context("spawn")
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, n, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    results = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    ranges = sorted(r[2] for r in results)
    assert ranges[0][0] == 0 and ranges[-1][1] == n and ranges[0][1] == ranges[1][0]
    for rank, full, _ in results:
        for k in ("global", "prob", "pred"):
            assert full[k].shape[0] == n
            assert torch.equal(full[k], seq[k]), (rank, k)      # gathered order == the reference's sequential order


def test_embedding_store_layout(tmp_path):
    emb = torch.randn(12, 128)
    labels = torch.randint(0, 2, (12, 5)).float()
    paths = save_embedding_chunks(emb, labels, str(tmp_path), chunk=5)
    assert [os.path.basename(p) for p in paths] == ["embeddings_dataset_5.pt", "embeddings_dataset_10.pt",
                                                    "embeddings_dataset_final.pt"]
    parts = [torch.load(p, weights_only=False) for p in paths]
    assert [len(d) for d in parts] == [5, 5, 2]
    glued = torch.utils.data.ConcatDataset(parts)                  # CSV_reformatting/glue_dataset.py:33-38
    assert torch.equal(torch.cat([d.tensors[0] for d in glued.datasets]), emb)     # Trainer.py:1252-1271 access pattern
    assert torch.equal(torch.cat([d.tensors[1] for d in glued.datasets]), labels)


def test_pack_and_unpack_rows_round_trip():
    from incremental_multimodal_medical_learning_ii_b200.extraction import pack_rows, unpack_rows
    g = torch.Generator().manual_seed(0)
    emb, prob = torch.randn(7, 128, generator=g), torch.rand(7, 14, generator=g)
    pred = (prob > 0.5).to(torch.uint8)
    packed = pack_rows([emb, prob, pred])
    assert packed.shape == (7, 128 * 4 + 14 * 4 + 14) and packed.dtype == torch.uint8       # 582 bytes per frame
    e2, p2, d2 = unpack_rows(packed, [((128,), torch.float32), ((14,), torch.float32), ((14,), torch.uint8)])
    assert torch.equal(e2, emb) and torch.equal(p2, prob) and torch.equal(d2, pred)
    empty = pack_rows([emb[:0], prob[:0], pred[:0]])
    assert empty.shape == (0, 582)
