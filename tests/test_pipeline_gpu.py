"""HostFramePipeline: double-buffered host->device feeding gives exactly the results of the synchronous loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_matches_synchronous_loop():
    import weights as Wt
    from incremental_multimodal_medical_learning_ii_b200 import frames as FR
    from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet
    from incremental_multimodal_medical_learning_ii_b200.pipeline import HostFramePipeline
    model = get_biovil_resnet(None)
    model.load_state_dict(Wt.make_state_dict(27, randomize_bn=True))
    model.eval().to("cuda:0")
    model.set_prompts(FR.synthetic_prompt_embeddings(14, 1, 128, seed=29), reduce="mean")
    batches = [FR.synthetic_frames_u8(8 * i, 8, 96, kind="structured", seed=0).pin_memory() for i in range(5)]
    ref = [{k: v.cpu() for k, v in model.embed_and_score(b.cuda()).items()} for b in batches]
    seen = []
    calls = []
    pipe = HostFramePipeline(model)
    for out in pipe.run(iter(batches), on_device_result=lambda r: calls.append(r["prob"].shape)):
        assert all(not t.is_cuda for t in out.values())
        seen.append({k: v.clone() for k, v in out.items()})      # host buffers are reused two batches later
    assert len(seen) == len(batches) == len(calls)
    for got, want in zip(seen, ref):
        for k in ("global", "prob", "pred"):
            assert torch.equal(got[k], want[k]), k


def test_pipeline_requires_cuda_model():
    from incremental_multimodal_medical_learning_ii_b200.image import get_biovil_resnet
    from incremental_multimodal_medical_learning_ii_b200.pipeline import HostFramePipeline
    with pytest.raises(RuntimeError):
        HostFramePipeline(get_biovil_resnet(None))
