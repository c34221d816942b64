"""Build + ctypes binding of the C-ABI library (``include/edtr_b200.h``).

The library is built in-tree (``edtr_b200/libedtr_b200.so``) with
``nvcc -gencode arch=compute_100a,code=sm_100a`` so that it travels with the repo
snapshot.  There is no CPU or PyTorch fallback: if the shared object cannot be
loaded, or the device is not sm_100-class, every op raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libedtr_b200.so")
SOURCES = ["api.cu", "gemm_conv.cu", "gemm2.cu", "attention.cu", "norm.cu", "elementwise.cu", "swin.cu"]
HEADER = os.path.join(_REPO_DIR, "include", "edtr_b200.h")

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
]

# every symbol include/edtr_b200.h declares
EXPORTED = [
    "edtr_last_error", "edtr_version", "edtr_set_device", "edtr_init", "edtr_set_workspace", "edtr_set_gemm_max_clusters", "edtr_gemm_tile_n", "edtr_gemm_bf16",
    "edtr_conv3x3_bf16", "edtr_conv3x3_up2x_bf16", "edtr_attention_bf16", "edtr_groupnorm_partial_size", "edtr_groupnorm_stats", "edtr_groupnorm_apply",
    "edtr_groupnorm_fused_supported", "edtr_groupnorm_fused", "edtr_groupnorm_pool", "edtr_groupnorm_apply_stats",
    "edtr_layernorm_bf16", "edtr_layernorm_padded_bf16", "edtr_pixel_unshuffle_f32_to_nhwc_bf16",
    "edtr_window_attention_bf16", "edtr_softmax_rows", "edtr_upsample2x_bf16", "edtr_im2col_bf16",
    "edtr_nchw_f32_to_nhwc_bf16", "edtr_pointwise_nchw_f32_to_nhwc_bf16", "edtr_nhwc_bf16_to_nchw", "edtr_cast_f32_to_bf16",
    "edtr_tile_blend", "edtr_timestep_embedding", "edtr_sampler_update", "edtr_wavelet_level",
]


class EdtrEpilogue(Structure):
    """Mirror of ``struct EdtrEpilogue`` (include/edtr_b200.h)."""

    _fields_ = [
        ("bias", c_void_p),
        ("rowvec", c_void_p),
        ("rowvec_ld", c_int32),
        ("rows_per_group", c_int32),
        ("residual", c_void_p),
        ("ldr", c_int32),
        ("out", c_void_p),
        ("ldc", c_int32),
        ("act", c_int32),
        ("out_mode", c_int32),
        ("hw", c_int32),
        ("alpha", c_float),
    ]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libedtr_b200.so")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``libedtr_b200.so`` (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    extra = os.environ.get("EDTR_NVCC_EXTRA", "").split()   # developer experiments (-DNAME=value)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB_PATH] + [os.path.join(CSRC_DIR, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lock = threading.Lock()
WORKSPACE_BYTES = 64 << 20
_workspace = None
_lib = None
_initialised = False


def _bind(lib: ctypes.CDLL) -> None:
    vp, ci = c_void_p, c_int
    ep = POINTER(EdtrEpilogue)
    lib.edtr_last_error.restype = c_char_p
    lib.edtr_last_error.argtypes = []
    lib.edtr_version.restype = ci
    lib.edtr_version.argtypes = []
    lib.edtr_set_device.restype = ci
    lib.edtr_set_device.argtypes = [ci]
    lib.edtr_init.restype = ci
    lib.edtr_set_workspace.restype = ci
    lib.edtr_set_workspace.argtypes = [vp, c_size_t]
    lib.edtr_init.argtypes = []
    lib.edtr_set_gemm_max_clusters.restype = ci
    lib.edtr_set_gemm_max_clusters.argtypes = [ci]
    lib.edtr_gemm_tile_n.restype = ci
    lib.edtr_gemm_tile_n.argtypes = [ci, ci, ci, ci]
    lib.edtr_gemm_bf16.restype = ci
    lib.edtr_gemm_bf16.argtypes = [vp, ci, vp, ci, ci, ci, ci, ep, vp]
    lib.edtr_conv3x3_bf16.restype = ci
    lib.edtr_conv3x3_bf16.argtypes = [vp, ci, ci, ci, ci, ci, vp, ci, ep, vp]
    lib.edtr_conv3x3_up2x_bf16.restype = ci
    lib.edtr_conv3x3_up2x_bf16.argtypes = [vp, ci, ci, ci, ci, ci, vp, ci, ep, vp]
    lib.edtr_attention_bf16.restype = ci
    lib.edtr_attention_bf16.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, c_float, vp]
    lib.edtr_groupnorm_partial_size.restype = c_size_t
    lib.edtr_groupnorm_partial_size.argtypes = [ci, ci, ci, ci]
    lib.edtr_groupnorm_stats.restype = ci
    lib.edtr_groupnorm_stats.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp]
    lib.edtr_groupnorm_fused_supported.restype = ci
    lib.edtr_groupnorm_fused_supported.argtypes = [ci, ci, ci, ci]
    lib.edtr_groupnorm_fused.restype = ci
    lib.edtr_groupnorm_fused.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, vp, vp, c_float, ci, vp]
    lib.edtr_wavelet_level.restype = ci
    lib.edtr_wavelet_level.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
    lib.edtr_groupnorm_pool.restype = ci
    lib.edtr_groupnorm_pool.argtypes = [vp, ci, ci, ci, ci, c_float, vp, vp]
    lib.edtr_groupnorm_apply_stats.restype = ci
    lib.edtr_groupnorm_apply_stats.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, vp, vp, vp, c_float, ci, vp]
    lib.edtr_groupnorm_apply.restype = ci
    lib.edtr_groupnorm_apply.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, vp, vp, vp, c_float, ci, vp]
    lib.edtr_layernorm_bf16.restype = ci
    lib.edtr_layernorm_bf16.argtypes = [vp, ci, vp, ci, ci, ci, vp, vp, c_float, vp]
    lib.edtr_layernorm_padded_bf16.restype = ci
    lib.edtr_layernorm_padded_bf16.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp, vp, c_float, vp]
    lib.edtr_pixel_unshuffle_f32_to_nhwc_bf16.restype = ci
    lib.edtr_pixel_unshuffle_f32_to_nhwc_bf16.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp, c_float, vp]
    lib.edtr_window_attention_bf16.restype = ci
    lib.edtr_window_attention_bf16.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, c_float, vp, vp, vp]
    lib.edtr_softmax_rows.restype = ci
    lib.edtr_softmax_rows.argtypes = [vp, ci, vp, ci, ci, ci, c_float, vp]
    lib.edtr_upsample2x_bf16.restype = ci
    lib.edtr_upsample2x_bf16.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, vp]
    lib.edtr_im2col_bf16.restype = ci
    lib.edtr_im2col_bf16.argtypes = [vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
    lib.edtr_nchw_f32_to_nhwc_bf16.restype = ci
    lib.edtr_nchw_f32_to_nhwc_bf16.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
    lib.edtr_pointwise_nchw_f32_to_nhwc_bf16.restype = ci
    lib.edtr_pointwise_nchw_f32_to_nhwc_bf16.argtypes = [vp, vp, vp, c_float, vp, ci, ci, ci, ci, ci, ci, vp]
    lib.edtr_nhwc_bf16_to_nchw.restype = ci
    lib.edtr_nhwc_bf16_to_nchw.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp]
    lib.edtr_cast_f32_to_bf16.restype = ci
    lib.edtr_cast_f32_to_bf16.argtypes = [vp, vp, c_size_t, vp]
    lib.edtr_tile_blend.restype = ci
    lib.edtr_tile_blend.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.edtr_timestep_embedding.restype = ci
    lib.edtr_timestep_embedding.argtypes = [vp, vp, ci, ci, c_float, vp]
    lib.edtr_sampler_update.restype = ci
    lib.edtr_sampler_update.argtypes = [vp] * 11 + [ci, ci, vp]


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building first if needed) the shared library. No device calls."""
    global _lib
    with _lock:
        if _lib is None:
            if build_if_missing and _stale():
                build()
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"{LIB_PATH} is missing and could not be built; no CPU fallback exists")
            lib = ctypes.CDLL(LIB_PATH)
            _bind(lib)
            _lib = lib
        return _lib


def use_workspace(index: int) -> None:
    """Select which of the two split-K scratch buffers later launches use (one per concurrent stream)."""
    lib = device_lib()
    check(lib.edtr_set_workspace(_workspace[index].data_ptr(), WORKSPACE_BYTES), "edtr_set_workspace")
    LAUNCHES[0] -= 1  # not a kernel launch


def set_gemm_max_clusters(n: int) -> None:
    """CTA pairs later GEMM / convolution launches may occupy (74 = the whole GPU)."""
    check(device_lib().edtr_set_gemm_max_clusters(int(n)), "edtr_set_gemm_max_clusters")
    LAUNCHES[0] -= 1  # not a kernel launch


def last_error() -> str:
    return load().edtr_last_error().decode("utf-8", "replace")


LAUNCHES = [0]  # successful kernel launches through the C-ABI


def check(rc: int, what: str) -> None:
    if rc == 0:
        LAUNCHES[0] += 1
        return
    msg = f"{what}: {last_error()} (code {rc})"
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(msg)


def device_lib() -> ctypes.CDLL:
    """Library handle for compute calls: requires a CUDA sm_100-class device."""
    global _initialised
    lib = load()
    if not _initialised:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("edtr_b200 has no CPU fallback: a CUDA (sm_100a) device is required")
        check(lib.edtr_set_device(torch.cuda.current_device()), "edtr_set_device")
        check(lib.edtr_init(), "edtr_init")
        # split-K scratch (stream-ordered use on the current stream; kept alive for the process lifetime)
        global _workspace
        _workspace = [torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device="cuda") for _ in range(2)]
        check(lib.edtr_set_workspace(_workspace[0].data_ptr(), WORKSPACE_BYTES), "edtr_set_workspace")
        _initialised = True
    return lib
