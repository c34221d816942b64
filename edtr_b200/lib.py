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
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libedtr_b200.so")
SOURCES = ["api.cu", "gemm_conv.cu", "gemm2.cu", "attention.cu", "norm.cu", "elementwise.cu", "swin.cu", "f32.cu"]
HEADER = os.path.join(_REPO_DIR, "include", "edtr_b200.h")

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
]

# every symbol include/edtr_b200.h declares
EXPORTED = [
    "edtr_last_error", "edtr_version", "edtr_set_device", "edtr_init", "edtr_gemm_workspace_size", "edtr_gemm_row_stats_parts", "edtr_gemm_tile_n", "edtr_gemm_bf16",
    "edtr_conv3x3_bf16", "edtr_conv3x3_up2x_bf16", "edtr_attention_bf16", "edtr_groupnorm_partial_size", "edtr_groupnorm_stats", "edtr_groupnorm_apply",
    "edtr_groupnorm_fused_supported", "edtr_groupnorm_fused", "edtr_groupnorm_pool", "edtr_groupnorm_apply_stats", "edtr_groupnorm_fold",
    "edtr_layernorm_bf16", "edtr_layernorm_padded_bf16", "edtr_pixel_unshuffle_f32_to_nhwc_bf16",
    "edtr_window_attention_bf16", "edtr_softmax_rows", "edtr_upsample2x_bf16", "edtr_im2col_bf16",
    "edtr_nchw_f32_to_nhwc_bf16", "edtr_pointwise_nchw_f32_to_nhwc_bf16", "edtr_nhwc_bf16_to_nchw", "edtr_cast_f32_to_bf16",
    "edtr_tile_blend", "edtr_timestep_embedding", "edtr_sampler_update", "edtr_wavelet_level",
    "edtr_f32_gemm", "edtr_f32_groupnorm_scratch_bytes", "edtr_f32_groupnorm", "edtr_f32_layernorm", "edtr_f32_softmax_rows",
    "edtr_f32_geglu", "edtr_f32_silu", "edtr_f32_nchw_to_nhwc", "edtr_f32_timestep_embedding",
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
        ("workspace", c_void_p),
        ("workspace_bytes", c_uint64),
        ("max_clusters", c_int32),
        ("ln_stats", c_void_p),
        ("ln_parts", c_int32),
        ("ln_c", c_int32),
        ("ln_eps", c_float),
        ("ln_colsum", c_void_p),
        ("row_stats", c_void_p),
        ("row_stats_cap", c_int32),
        ("gn_partial", c_void_p),
        ("gn_hw", c_int32),
        ("gn_slabs", c_int32),
        ("gn_slab0", c_int32),
        ("gn_unit", c_int32),
    ]


class EdtrF32Gemm(Structure):
    """Mirror of ``struct EdtrF32Gemm`` (include/edtr_b200.h): the generic fp32 implicit GEMM of the fp32 mode."""

    _fields_ = [
        ("A", c_void_p), ("W", c_void_p), ("C", c_void_p),
        ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("lda", c_int64), ("ldw", c_int64), ("ldc", c_int64),
        ("w_kn", c_int32), ("batch1", c_int32), ("batch2", c_int32),
        ("a_stride1", c_int64), ("a_stride2", c_int64), ("w_stride1", c_int64), ("w_stride2", c_int64),
        ("c_stride1", c_int64), ("c_stride2", c_int64),
        ("conv", c_int32), ("H", c_int32), ("W_in", c_int32), ("Cin", c_int32), ("Ho", c_int32), ("Wo", c_int32),
        ("conv_stride", c_int32), ("pad_top", c_int32), ("pad_left", c_int32), ("up2x", c_int32),
        ("alpha", c_float), ("bias", c_void_p), ("rowvec", c_void_p), ("rowvec_ld", c_int64),
        ("rows_per_group", c_int32), ("residual", c_void_p), ("ldr", c_int64),
        ("act", c_int32), ("out_nchw", c_int32), ("hw", c_int32),
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
WORKSPACE_BYTES = 64 << 20      # split-K scratch (fp32 partial tiles) per (device, slot)
_lib = None
_torch = None
_devices = {}                   # device index -> {"workspaces": {slot: tensor}}: per-device init, no global device state
_tls = threading.local()        # per host thread: workspace slot and CTA-pair share of the launches issued from it


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
    lib.edtr_init.argtypes = []
    lib.edtr_gemm_workspace_size.restype = c_size_t
    lib.edtr_gemm_workspace_size.argtypes = [ci, ci, ci]
    lib.edtr_gemm_row_stats_parts.restype = ci
    lib.edtr_gemm_row_stats_parts.argtypes = [ci, ci, ci, ep]
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
    cll = ctypes.c_longlong
    lib.edtr_f32_gemm.restype = ci
    lib.edtr_f32_gemm.argtypes = [POINTER(EdtrF32Gemm), vp]
    lib.edtr_f32_groupnorm_scratch_bytes.restype = c_size_t
    lib.edtr_f32_groupnorm_scratch_bytes.argtypes = [ci, ci, ci]
    lib.edtr_f32_groupnorm.restype = ci
    lib.edtr_f32_groupnorm.argtypes = [vp, cll, vp, cll, ci, ci, ci, ci, vp, vp, c_float, ci, vp, vp]
    lib.edtr_f32_layernorm.restype = ci
    lib.edtr_f32_layernorm.argtypes = [vp, cll, vp, cll, ci, ci, vp, vp, c_float, vp]
    lib.edtr_f32_softmax_rows.restype = ci
    lib.edtr_f32_softmax_rows.argtypes = [vp, cll, cll, ci, c_float, vp]
    lib.edtr_f32_geglu.restype = ci
    lib.edtr_f32_geglu.argtypes = [vp, cll, vp, cll, cll, ci, vp]
    lib.edtr_f32_silu.restype = ci
    lib.edtr_f32_silu.argtypes = [vp, vp, cll, vp]
    lib.edtr_f32_nchw_to_nhwc.restype = ci
    lib.edtr_f32_nchw_to_nhwc.argtypes = [vp, vp, cll, ci, ci, ci, ci, c_float, vp]
    lib.edtr_f32_timestep_embedding.restype = ci
    lib.edtr_f32_timestep_embedding.argtypes = [vp, vp, ci, ci, c_float, vp]
    lib.edtr_groupnorm_fold.restype = ci
    lib.edtr_groupnorm_fold.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp]
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


def use_workspace(index) -> None:
    """Which split-K scratch slot launches issued by THIS host thread use from now on.  A slot is any hashable name; the
    engines use (workspace uid, 0 | 1): one slot per set of static buffers and per branch that may run concurrently
    (0 main stream, 1 the ControlNet side stream), so batches in flight on different streams never share scratch.
    Python-side selection only: the buffer is passed to the library with every call (EdtrEpilogue.workspace); the
    library itself keeps no state."""
    _tls.slot = index


def set_gemm_max_clusters(n: int) -> None:
    """CTA pairs later GEMM / convolution launches of this host thread may occupy (74 = the whole GPU); passed per
    call (EdtrEpilogue.max_clusters)."""
    n = int(n)
    if not 1 <= n <= 74:
        raise ValueError(f"clusters must be in [1, 74], got {n}")
    _tls.max_clusters = n


def gemm_scratch(device_index: int):
    """(pointer, bytes, max_clusters) of the current thread's split-K workspace on `device_index`."""
    st = _devices[device_index]
    slot = getattr(_tls, "slot", 0)
    ws = st["workspaces"].get(slot)
    if ws is None:
        import torch

        ws = st["workspaces"][slot] = torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=f"cuda:{device_index}")
    return ws.data_ptr(), WORKSPACE_BYTES, getattr(_tls, "max_clusters", 74)


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


def device_lib(device_index=None) -> ctypes.CDLL:
    """Library handle for compute calls on `device_index` (default: torch's current device): requires a CUDA
    sm_100-class device; kernel attributes are primed once per device.  The caller must have made the device current
    (ops does so with torch.cuda.device(...) around every launch sequence)."""
    global _torch
    if _torch is None:
        import torch as _t

        _torch = _t
    torch = _torch
    if device_index is None:
        device_index = torch.cuda.current_device() if torch.cuda.is_initialized() else None
    if device_index is not None and device_index in _devices:      # hot path: a few lookups per launch
        if getattr(_tls, "device", None) != device_index:
            # The library links its own (static) CUDA runtime, whose per-thread current device is independent of
            # torch's: keep it on the device the launches are meant for (a NULL / default stream handle is valid on
            # every device, so a mismatch would silently run the kernel on the wrong GPU).
            check(_lib.edtr_set_device(device_index), "edtr_set_device")
            LAUNCHES[0] -= 1
            _tls.device = device_index
        return _lib
    lib = load()
    if not torch.cuda.is_available():
        raise RuntimeError("edtr_b200 has no CPU fallback: a CUDA (sm_100a) device is required")
    if device_index is None:
        device_index = torch.cuda.current_device()
    if device_index not in _devices:
        with _lock:
            if device_index not in _devices:
                with torch.cuda.device(device_index):
                    check(lib.edtr_set_device(device_index), "edtr_set_device")
                    LAUNCHES[0] -= 1
                    check(lib.edtr_init(), "edtr_init")
                    LAUNCHES[0] -= 1
                _devices[device_index] = {"workspaces": {}}
    if getattr(_tls, "device", None) != device_index:
        check(lib.edtr_set_device(device_index), "edtr_set_device")
        LAUNCHES[0] -= 1
        _tls.device = device_index
    return lib
