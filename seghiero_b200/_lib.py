"""ctypes binding of libseghiero_b200.so (the C ABI declared in include/seghiero_b200.h).

There is no CPU fallback: if the library is missing or a call fails, the op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libseghiero_b200.so")

_p, _i, _l, _f, _d, _sz = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/seghiero_b200.h one to one
SIGNATURES = {
    "sh_targets_two_level": (_i, [_p, _i, _p, _l, _p, _i, _p]),
    "sh_targets_three_level": (_i, [_p, _i, _p, _p, _l, _p, _p, _i, _p, _p]),
    "sh_targets_gather": (_i, [_p, _i, _p, _l, _p, _i, _p, _p]),
    "sh_colorize": (_i, [_p, _i, _l, _p, _i, _p, _p, _p]),
    "sh_decode": (_i, [_p, _i, _i, _i, _l, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p]),
    "sh_bce2_grid": (_i, [_i, _l, _i, _i]),
    "sh_bce2_fwdbwd": (_i, [_p, _i, _p, _i, _p, _i, _l, _i, _i, _p, _i, _i, _f, _f, _p, _p, _p, _p, _i, _p]),
    "sh_loss2_final": (_i, [_p, _p, _i, _i, _d, _p, _d, _p, _p, _f, _p, _p]),
    "sh_scale_inplace": (_i, [_p, _i, _l, _p, _p]),
    "sh_rmi3_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "sh_rmi3_workspace_offsets": (_i, [_i, _i, _i, _i, _i, _i, _p]),
    "sh_rmi3_fast_path": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i]),
    "sh_rmi3_pass2_kind": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i]),
    "sh_rmi3_forward": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i, _f, _f, _p, _i, _p]),
    "sh_loss3_final": (_i, [_i, _i, _i, _i, _i, _i, _p, _f, _p, _d, _p, _p, _f, _p, _p]),
    "sh_rmi3_backward": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _f, _p, _p, _i, _p]),
    "sh_triplet_forward": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "sh_triplet_backward": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "sh_upsample_bilinear": (_i, [_p, _i, _p, _l, _i, _i, _i, _i, _p]),
    "sh_upsample_bilinear_adjoint": (_i, [_p, _i, _p, _l, _i, _i, _i, _i, _p]),
    "sh_aux_ce_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "sh_aux_ce_fwdbwd": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "sh_decode_upsampled": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _p, _p]),
}

_lib = None


class SegHieroLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SegHieroLibraryError(
            f"{LIB_PATH} not found: build it with `python -m seghiero_b200.build` "
            "(seghiero_b200 has no CPU or eager fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError -> missing export
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, name: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        what = {-1: "bad argument", -2: "unsupported configuration"}.get(rc, "error")
        raise SegHieroLibraryError(f"{name}: {what} (code {rc})")
    raise SegHieroLibraryError(f"{name}: CUDA launch failed with cudaError {rc}")


def call(name: str, *args):
    rc = getattr(load(), name)(*args)
    check(rc, name)
