"""ctypes binding of the C-ABI library (include/vaesne_b200.h).

The product path has exactly one backend: ``lib/libvaesne_b200.so`` built by nvcc for sm_100a
(``python __graft_entry__.py`` / ``vaesne-dev_b200/build.py``).  If it is missing, or a tensor is
not on a CUDA device, the ops raise — there is no CPU fallback.  The test-suite may point the
loader at the CUDA-semantics emulator build of the same sources (tests/emu) through
``use_library(path)``; that hook is never taken implicitly.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libvaesne_b200.so")

_lib = None
_emulated = False

_vp, _i, _ll, _f, _u32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint32

_SIGS = {
    "vaesne_lin_fwd": [_vp, _ll, _vp, _ll, _i, _i, _i, _vp, _vp, _i, _vp, _ll, _vp, _ll, _vp, _vp, _f, _vp, _f, _vp, _u32, _vp, _ll, _vp],
    "vaesne_lin_bwd": [_vp, _ll, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _ll, _i, _f, _vp, _u32, _i, _vp, _ll,
                       _vp, _ll, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _i, _vp],
    "vaesne_attn_fwd": [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _i, _i, _f, _vp, _u32, _vp, _ll, _vp, _vp],
    "vaesne_attn_bwd": [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _i, _i, _f, _vp, _u32, _vp, _ll, _vp, _vp, _ll,
                        _vp, _vp, _ll, _vp, _ll, _vp, _ll, _vp],
    "vaesne_attn_fwd_ex": [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _i, _i, _f, _vp, _u32, _vp, _ll, _vp, _i, _vp],
    "vaesne_attn_bwd_ex": [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _vp, _i, _i, _f, _vp, _u32, _vp, _ll, _vp, _vp, _ll,
                           _vp, _vp, _ll, _vp, _ll, _vp, _ll, _i, _vp],
    "vaesne_attn_combine": [_vp, _vp, _i, _ll, _i, _i, _i, _vp, _ll, _vp, _vp],
    "vaesne_sincos_feat": [_vp, _ll, _vp, _i, _vp, _ll, _vp],
    "vaesne_gather_rows": [_vp, _ll, _vp, _i, _vp, _ll, _i, _vp],
    "vaesne_scatter_rows": [_vp, _ll, _vp, _ll, _vp, _i, _vp],
    "vaesne_expand_rows": [_vp, _ll, _ll, _i, _vp, _vp],
    "vaesne_expand_rows_bwd": [_vp, _ll, _ll, _i, _vp, _i, _vp],
    "vaesne_copy3d": [_vp, _ll, _ll, _vp, _ll, _ll, _ll, _ll, _ll, _i, _vp],
    "vaesne_latent_fwd": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "vaesne_latent_bwd": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp],
    "vaesne_kl_fwd": [_vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp],
    "vaesne_loglik_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp, _i, _vp],
    "vaesne_loglik_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp, _f, _vp, _vp, _vp],
    "vaesne_kl_bwd": [_vp, _vp, _i, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp],
    "vaesne_scale": [_vp, _ll, _f, _vp, _vp, _vp],
    "vaesne_iwae_combine": [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "vaesne_elbo_combine": [_vp, _vp, _i, _i, _vp, _vp],
    "vaesne_adamw_flat": [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _vp, _f, _vp],
    "vaesne_step_advance": [_vp, _vp, _vp],
    "vaesne_seed_next": [_vp, _vp, _vp],
    "vaesne_l2norm_fwd": [_vp, _i, _i, _f, _vp, _vp, _vp],
    "vaesne_l2norm_bwd": [_vp, _vp, _vp, _i, _i, _vp, _i, _vp],
    "vaesne_ce_rows_fwd": [_vp, _i, _vp, _i, _i, _f, _i, _vp, _vp, _vp],
    "vaesne_ce_rows_bwd": [_vp, _i, _vp, _i, _i, _f, _i, _vp, _f, _vp, _vp, _i, _vp, _i, _vp],
    "vaesne_sum_scale": [_vp, _vp, _i, _f, _vp, _vp],
    "vaesne_augment": [_vp, _vp, _ll, _ll, _i, _f, _f, _f, _vp, _u32, _vp, _vp, _vp],
}
EXPORTS = sorted(list(_SIGS) + ["vaesne_last_error", "vaesne_abi_version", "vaesne_is_emulated", "vaesne_launch_count",
                  "vaesne_debug_tc", "vaesne_debug_tc_prof"])      # the last two: probe hooks (tests/probe)


def _bind(path: str):
    lib = C.CDLL(path)
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = sig
        fn.restype = C.c_int
    lib.vaesne_last_error.restype = C.c_char_p
    lib.vaesne_abi_version.restype = C.c_int
    lib.vaesne_is_emulated.restype = C.c_int
    lib.vaesne_launch_count.restype = C.c_longlong
    return lib


def use_library(path: str) -> None:
    """Explicitly select the shared library (tests use this for the emulator build)."""
    global _lib, _emulated
    _lib = _bind(path)
    _emulated = bool(_lib.vaesne_is_emulated())


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"VAESNe-B200: CUDA library not built ({LIB_PATH}). Run `python __graft_entry__.py` "
                "(or vaesne-dev_b200/build.py); there is no CPU fallback.")
        use_library(LIB_PATH)
    return _lib


def launch_count() -> int:
    return int(lib().vaesne_launch_count())


def is_emulated() -> bool:
    lib()
    return _emulated


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"VAESNe-B200 kernel error {rc}: {lib().vaesne_last_error().decode()}")


def stream_of(t: torch.Tensor) -> int:
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    if not is_emulated():
        raise RuntimeError("VAESNe-B200 ops need CUDA tensors (no CPU fallback); got a CPU tensor")
    return 0


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def ptr_table(tensors):
    """Host array of device pointers (kept alive by the caller for the duration of the call)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = 0 if t is None else t.data_ptr()
    return arr


def int_table(vals):
    arr = (C.c_int * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr
