"""ctypes binding of ``libbiovil_b200.so`` (the C ABI in ``include/biovil_b200.h``).

There is deliberately no fallback: if the shared library is missing, or the machine has no sm_100 GPU, the
product path raises.  ``build()`` compiles the library in-tree with nvcc (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_size_t, c_uint8, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libbiovil_b200.so"
CSRC = PKG_DIR / "csrc"
INCLUDE = REPO_ROOT / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]

BV_DTYPE_U8 = 0
BV_DTYPE_F32 = 1
BV_NUM_BLOCKS = 16

BV_OK, BV_ERR_INVALID, BV_ERR_NO_DEVICE, BV_ERR_CUDA, BV_ERR_WORKSPACE = 0, -1, -2, -3, -4


class NativeError(RuntimeError):
    """A call into libbiovil_b200 failed (``.code`` holds the bv_status)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"biovil_b200 error {code}: {message}")
        self.code = code


class BvConv(Structure):
    _fields_ = [("w", c_void_p), ("bias", c_void_p), ("cin", c_int32), ("cout", c_int32), ("r", c_int32),
                ("s", c_int32), ("stride", c_int32), ("pad", c_int32)]


class BvWeights(Structure):
    _fields_ = [("stem_u8", BvConv), ("stem_f1", BvConv), ("stem_f3", BvConv),
                ("conv1", BvConv * BV_NUM_BLOCKS), ("conv2", BvConv * BV_NUM_BLOCKS),
                ("conv3", BvConv * BV_NUM_BLOCKS), ("downsample", BvConv * BV_NUM_BLOCKS),
                ("proj0", BvConv), ("proj3_wt", c_void_p), ("proj3_b", c_void_p), ("stem_u8_k8", BvConv)]


class BvOutputs(Structure):
    _fields_ = [("global_emb", c_void_p), ("patch_emb", c_void_p), ("normalize_patch", c_int32),
                ("pooled", c_void_p), ("trunk_nhwc_bf16", c_void_p), ("sim", c_void_p), ("prob", c_void_p),
                ("pred", c_void_p), ("score", c_void_p), ("heat", c_void_p)]


class BvLaunchInfo(Structure):
    _fields_ = [("name", ctypes.c_char * 64), ("flops", ctypes.c_double), ("bytes", ctypes.c_double),
                ("ms", c_float)]


# Every symbol include/biovil_b200.h declares (tests check the library exports all of them).
EXPORTED_SYMBOLS = (
    "bv_last_error", "bv_version", "bv_workspace_bytes", "bv_patch_grid", "bv_create", "bv_destroy",
    "bv_set_prompts", "bv_forward", "bv_score", "bv_last_forward_launches", "bv_set_profile", "bv_get_profile",
    "bv_conv2d_nhwc", "bv_conv_chain_nhwc", "bv_smooth_heatmaps",
    "bv_resize_workspace_bytes", "bv_resize_center_crop_u8", "bv_pair_gemm_test", "bv_l1_block_nhwc",
    "bv_stem_u8_nhwc", "bv_stem_conv1_u8_nhwc", "bv_forward_graph", "bv_pairwise_cosine", "bv_quantize_frames_f32",
    "bv_jpeg_info", "bv_jpeg_decode_gray_u8", "bv_l1_block_ds_nhwc", "bv_pair_chain_nhwc", "bv_jpeg_decode_batch_gray_u8",
    "bv_heatmaps_to_image_size",
)

_lib = None


def sources():
    return [CSRC / "biovil_b200.cu"]


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile ``libbiovil_b200.so`` for sm_100a next to this file (no-op when up to date)."""
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [INCLUDE / "biovil_b200.h"]
    if LIB_PATH.exists() and not force:
        newest = max(p.stat().st_mtime for p in deps)
        if LIB_PATH.stat().st_mtime >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, sources()), "-ldl"]
    if os.environ.get("BV_BUILD_TIMING"):          # profiling build: wait-cycle counters compiled in (BV_TIMING=1 reports)
        cmd.insert(1, "-DBV_ENABLE_TIMING=1")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the shared library (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("BV_LIB_PATH", str(LIB_PATH)))      # A/B of two builds on one box (tools/gpu_ab.sh)
    if not path.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the B200 path has no CPU / eager fallback)")
    l = ctypes.CDLL(str(path))
    if "BV_LIB_PATH" in os.environ:
        # an older build may lack entry points added since: give them a stub so the bindings below still resolve
        class _Tolerant:
            def __init__(self, lib):
                object.__setattr__(self, "_lib", lib)

            def __getattr__(self, name):
                try:
                    return getattr(self._lib, name)
                except AttributeError:
                    return type("_Missing", (), {"restype": None, "argtypes": None})()
        l = _Tolerant(l)
    l.bv_last_error.restype = c_char_p
    l.bv_version.restype = c_char_p
    l.bv_workspace_bytes.restype = c_size_t
    l.bv_workspace_bytes.argtypes = [c_int32] * 4
    l.bv_patch_grid.restype = c_int32
    l.bv_patch_grid.argtypes = [c_int32]
    l.bv_create.restype = c_int32
    l.bv_create.argtypes = [POINTER(c_void_p), POINTER(BvWeights), c_int32]
    l.bv_destroy.restype = None
    l.bv_destroy.argtypes = [c_void_p]
    l.bv_set_prompts.restype = c_int32
    l.bv_set_prompts.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    l.bv_forward.restype = c_int32
    l.bv_forward.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_size_t,
                             POINTER(BvOutputs), c_void_p]
    l.bv_score.restype = c_int32
    l.bv_score.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    l.bv_last_forward_launches.restype = c_int32
    l.bv_last_forward_launches.argtypes = [c_void_p]
    l.bv_set_profile.restype = c_int32
    l.bv_set_profile.argtypes = [c_void_p, c_int32]
    l.bv_get_profile.restype = c_int32
    l.bv_get_profile.argtypes = [c_void_p, POINTER(BvLaunchInfo), c_int32]
    l.bv_conv2d_nhwc.restype = c_int32
    l.bv_conv2d_nhwc.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(BvConv), c_void_p, c_int32, c_int32,
                                 POINTER(BvConv), c_void_p, c_int32, c_void_p, c_int32, c_void_p]
    l.bv_resize_workspace_bytes.restype = c_size_t
    l.bv_resize_workspace_bytes.argtypes = [c_int32] * 5
    l.bv_resize_center_crop_u8.restype = c_int32
    l.bv_resize_center_crop_u8.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                           c_size_t, c_void_p]
    l.bv_l1_block_nhwc.restype = c_int32
    l.bv_l1_block_nhwc.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(BvConv), POINTER(BvConv), c_void_p, c_void_p,
                                   POINTER(BvConv), c_void_p, c_void_p]
    l.bv_l1_block_ds_nhwc.restype = c_int32
    l.bv_l1_block_ds_nhwc.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(BvConv), POINTER(BvConv), c_void_p,
                                      POINTER(BvConv), c_void_p, POINTER(BvConv), c_void_p, c_void_p]
    l.bv_pair_chain_nhwc.restype = c_int32
    l.bv_pair_chain_nhwc.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(BvConv), c_void_p, c_void_p,
                                     POINTER(BvConv), c_void_p, c_void_p]
    l.bv_stem_u8_nhwc.restype = c_int32
    l.bv_stem_u8_nhwc.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(BvConv), c_void_p, c_int32, c_void_p]
    l.bv_stem_conv1_u8_nhwc.restype = c_int32
    l.bv_stem_conv1_u8_nhwc.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(BvConv), POINTER(BvConv), c_void_p,
                                        c_void_p, c_void_p]
    l.bv_pair_gemm_test.restype = c_int32
    l.bv_pair_gemm_test.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    l.bv_smooth_heatmaps.restype = c_int32
    l.bv_smooth_heatmaps.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p]
    l.bv_heatmaps_to_image_size.restype = c_int32
    l.bv_heatmaps_to_image_size.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                            c_void_p, c_void_p]
    l.bv_conv_chain_nhwc.restype = c_int32
    l.bv_conv_chain_nhwc.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(BvConv), c_void_p, c_int32, c_int32,
                                     POINTER(BvConv), c_void_p, c_void_p, POINTER(BvConv), c_void_p, c_void_p]
    l.bv_forward_graph.restype = c_int32
    l.bv_forward_graph.argtypes = l.bv_forward.argtypes
    l.bv_pairwise_cosine.restype = c_int32
    l.bv_pairwise_cosine.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    l.bv_quantize_frames_f32.restype = c_int32
    l.bv_quantize_frames_f32.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    l.bv_jpeg_info.restype = c_int32
    l.bv_jpeg_info.argtypes = [c_void_p, c_size_t, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]
    l.bv_jpeg_decode_batch_gray_u8.restype = c_int32
    l.bv_jpeg_decode_batch_gray_u8.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p]
    l.bv_jpeg_decode_gray_u8.restype = c_int32
    l.bv_jpeg_decode_gray_u8.argtypes = [c_void_p, c_size_t, c_void_p, c_int32, c_int32, c_int32, c_void_p]
    _lib = l
    return l


def check(code: int) -> None:
    if code != 0:
        raise NativeError(code, lib().bv_last_error().decode("utf-8", "replace"))


def ptr(t) -> c_void_p:
    """Device (or host) address of a torch tensor as a ctypes void pointer; ``None`` -> NULL."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def current_stream_handle(device) -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
