"""Host-side behaviour of the drop-in Python surface (no GPU): names, state_dict compatibility, error conventions."""
import os

import pytest
import torch

import weights as Wt


def test_health_multimodal_alias_exports_reference_names():
    import health_multimodal.image as hi
    from health_multimodal.image import ImageInferenceEngine, ImageModel, ResnetType, get_biovil_resnet  # noqa: F401
    from health_multimodal.image import get_biovil_resnet_inference  # noqa: F401
    from health_multimodal.image.model.model import JOINT_FEATURE_SIZE, MODEL_TYPE, ImageModelOutput  # noqa: F401
    from health_multimodal.image.data.transforms import (ExpandChannels, create_chest_xray_transform_for_inference,
                                                         infer_resize_params)  # noqa: F401
    from health_multimodal.image.utils import TRANSFORM_CENTER_CROP_SIZE, TRANSFORM_RESIZE
    assert (TRANSFORM_RESIZE, TRANSFORM_CENTER_CROP_SIZE) == (512, 480)
    assert MODEL_TYPE == "resnet50" and JOINT_FEATURE_SIZE == 128
    assert set(hi.__all__) >= {"ImageModel", "ResnetType", "ImageInferenceEngine", "get_biovil_resnet",
                               "get_biovil_resnet_inference"}


def test_state_dict_is_key_compatible_and_round_trips(tmp_path):
    from health_multimodal.image import get_biovil_resnet
    m = get_biovil_resnet(None)
    assert m.training                                  # the reference leaves the module in train mode (model.py:112)
    assert m.feature_size == 2048 and m.classifier is None and m.freeze_encoder is False
    sd_ref = Wt.make_state_dict(27, randomize_bn=True)  # keys/shapes verified against the reference in make_golden.py
    sd = m.state_dict()
    assert len(sd) == 328 and list(sd.keys()) == list(sd_ref.keys()) or set(sd) == set(sd_ref)
    for k, v in sd_ref.items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    m.load_state_dict(sd_ref)                           # strict
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd_ref[k])
    path = tmp_path / "biovil_image_resnet50_proj_size_128.pt"
    torch.save(sd_ref, path)
    m2 = get_biovil_resnet(str(path))                   # chexpert-get-embedding.py:38-40 style
    assert torch.equal(m2.state_dict()["projector.model.3.bias"], sd_ref["projector.model.3.bias"])
    m3 = get_biovil_resnet(path)                        # pathlib.Path accepted too
    assert torch.equal(m3.state_dict()["encoder.encoder.conv1.weight"], sd_ref["encoder.encoder.conv1.weight"])
    with pytest.raises(TypeError):
        get_biovil_resnet(123)                          # model.py:115-116
    bad = dict(sd_ref)
    bad.pop("encoder.encoder.fc.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)


def test_train_signature_and_guards():
    from health_multimodal.image import ImageModel, get_biovil_resnet
    m = get_biovil_resnet(None)
    assert m.train(mode=False, my_freeze=True) is m and not m.training
    m.train(True, my_freeze=True)
    assert m.training and not m.encoder.training and not m.projector.training
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))                    # CPU tensors: the product path must fail loudly
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(AssertionError):
        m.get_patchwise_projected_embeddings(torch.zeros(1, 3, 64, 64), normalize=True)   # model.py:169
    with pytest.raises(NotImplementedError):
        ImageModel("resnet18", 128)
    with pytest.raises(NotImplementedError):
        ImageModel("vgg", 128)                          # model.py:192
    with pytest.raises(RuntimeError):
        m.encoder.encoder.layer1[0](torch.zeros(1, 64, 8, 8))     # parameter containers have no eager forward


def test_engine_and_transforms():
    from torchvision.transforms import CenterCrop, Compose, Normalize, Resize, ToTensor
    from health_multimodal.image import ImageInferenceEngine, get_biovil_resnet
    from health_multimodal.image.data.transforms import (ExpandChannels, create_chest_xray_transform_for_inference,
                                                         infer_resize_params)
    t = create_chest_xray_transform_for_inference(512, 480)
    assert infer_resize_params(t.transforms) == (512, 480)
    with pytest.raises(ValueError):
        infer_resize_params([Normalize(0, 1)])
    with pytest.raises(ValueError):
        infer_resize_params([CenterCrop(480), Resize(512)])
    with pytest.raises(ValueError):
        ExpandChannels()(torch.zeros(3, 4, 4))
    x = torch.rand(1, 5, 7)
    y = ExpandChannels()(x)
    assert y.shape == (3, 5, 7) and torch.equal(y[0], y[2])
    m = get_biovil_resnet(None)
    eng = ImageInferenceEngine(m, t)
    assert not m.training and (eng.resize_size, eng.crop_size) == (512, 480) and eng.to == m.to
    with pytest.raises(AssertionError):
        ImageInferenceEngine(torch.nn.Linear(1, 1), t)
    # the PIL pipeline yields exactly k/255 with three identical channels: the contract the u8 fast path relies on
    from PIL import Image
    import numpy as np
    img = Image.fromarray((np.arange(600 * 520) % 251).astype(np.uint8).reshape(600, 520))
    out = t(img)
    assert out.shape == (3, 480, 480) and torch.equal(out[0], out[1])
    k = torch.round(out[0] * 255)
    assert (out[0] * 255 - k).abs().max() < 1e-3


def test_bn_fold_and_stem_packing_match_conv_bn():
    import torch.nn.functional as F
    from incremental_multimodal_medical_learning_ii_b200.packing import BN_EPS, block_names, fold_bn
    sd = Wt.make_state_dict(27, randomize_bn=True)
    names = block_names()
    assert len(names) == 16 and [s for _, s in names].count(2) == 3
    p = "encoder.encoder.layer2.0"
    x = torch.randn(2, 256, 10, 10, dtype=torch.float64)
    w, b = fold_bn(sd[p + ".downsample.0.weight"], sd, p + ".downsample.1")
    ref = F.batch_norm(F.conv2d(x, sd[p + ".downsample.0.weight"].double(), stride=2),
                       sd[p + ".downsample.1.running_mean"].double(), sd[p + ".downsample.1.running_var"].double(),
                       sd[p + ".downsample.1.weight"].double(), sd[p + ".downsample.1.bias"].double(),
                       training=False, eps=BN_EPS)
    assert torch.allclose(F.conv2d(x, w, b, stride=2), ref, atol=1e-10)
    # stem: three identical channels fold into one; 1/255 folds into the weights
    k = torch.randint(0, 256, (1, 1, 20, 20)).double()
    w, b = fold_bn(sd["encoder.encoder.conv1.weight"], sd, "encoder.encoder.bn1")
    full = F.conv2d((k / 255).repeat(1, 3, 1, 1), w, b, stride=2, padding=3)
    folded = F.conv2d(k, w.sum(1, keepdim=True) / 255, b, stride=2, padding=3)
    assert torch.allclose(full, folded, atol=1e-10)


def test_alias_package_leaves_the_reference_text_side_reachable():
    """The reference's scripts import ``health_multimodal.text`` next to ``health_multimodal.image``
    (Trainer.py, ZERO_JOINT_BOUNDS.py ...).  With this repository ahead of the reference on sys.path the image side must
    be ours and the text side must still resolve to the reference's package (skipped where no checkout is present)."""
    import importlib
    import os
    import subprocess
    import sys
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "health_multimodal", "text")):
        pytest.skip("no reference checkout on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import health_multimodal, health_multimodal.image as i, importlib.util as u; "
            "s = u.find_spec('health_multimodal.text'); "
            "print(i.__name__); print(s.origin if s else None)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([root, ref]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "incremental_multimodal_medical_learning_ii_b200.image"
    assert lines[1].startswith(os.path.join(ref, "health_multimodal", "text"))
