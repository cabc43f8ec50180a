"""ctypes binding of ``libdafne_b200.so`` (C ABI declared in ``include/dafne_b200.h``).

The product path has no CPU or eager fallback: if the shared library is missing or fails to load, importing a
symbol from here raises immediately with the build command to run.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdafne_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

DET_STRIDE = 20
MAX_LEVELS = 5


class DafneError(RuntimeError):
    pass


class ModelSpecC(C.Structure):
    _fields_ = [
        ("resnet_depth", C.c_int32),
        ("num_classes", C.c_int32),
        ("sort_corners", C.c_int32),
        ("thresh_with_ctr", C.c_int32),
        ("pre_nms_topk", C.c_int32),
        ("post_nms_topk", C.c_int32),
        ("score_thresh", C.c_float),
        ("nms_thresh", C.c_float),
        ("num_levels", C.c_int32),
        ("fpn_strides", C.c_int32 * MAX_LEVELS),
        ("pixel_mean", C.c_float * 3),
        ("pixel_std", C.c_float * 3),
        ("vehicle_merge", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class OpProfileC(C.Structure):
    _fields_ = [
        ("ms", C.c_float),
        ("kind", C.c_int32), ("block_n", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32), ("hout", C.c_int32), ("wout", C.c_int32),
        ("flops", C.c_double), ("bytes", C.c_double),
        ("name", C.c_char * 48),
    ]


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise DafneError("building libdafne_b200.so failed (make -C dafne_b200/csrc)")
    return LIB_PATH


_lib = None

_vp, _i, _f, _fp, _ip = C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p

_SIGNATURES = {
    "dafne_last_error": (C.c_char_p, []),
    "dafne_abi_version": (_i, []),
    "dafne_ctx_create": (_i, [C.POINTER(ModelSpecC), _i, C.POINTER(_vp)]),
    "dafne_ctx_destroy": (None, [_vp]),
    "dafne_load_weights": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(C.c_int64), _vp]),
    "dafne_weights_finalize": (_i, [_vp, _vp]),
    "dafne_workspace_bytes": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_size_t)]),
    "dafne_bind_workspace": (_i, [_vp, _i, _i, _i, _vp, C.c_size_t]),
    "dafne_forward_dense": (_i, [_vp, _vp, _i, C.POINTER(C.c_int32), _vp]),
    "dafne_head_output": (_i, [_vp, _i, _i, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "dafne_postprocess": (_i, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i, _vp, _vp, _i, _vp]),
    "dafne_postprocess_external": (
        _i,
        [_vp, _i, C.POINTER(C.c_int32), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int32),
         C.POINTER(C.c_int32), _i, _vp, _vp, _i, _vp, C.c_size_t, _vp],
    ),
    "dafne_postprocess_scratch_bytes": (_i, [_vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]),
    "dafne_detect": (_i, [_vp, _vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp, _vp, _i, _vp]),
    "dafne_graph_capture": (_i, [_vp, _vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i, _vp, _vp, _i, _vp]),
    "dafne_graph_launch": (_i, [_vp, _vp]),
    "dafne_detect_host": (_i, [_vp, _vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp, _vp, _i, _vp]),
    "dafne_detect_host_begin": (_i, [_vp, _vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _vp, _vp, _i, _vp,
                                     C.POINTER(_i)]),
    "dafne_detect_host_end": (_i, [_vp, _i]),
    "dafne_host_slot_wire": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(C.c_size_t), C.POINTER(_i)]),
    "dafne_voc_match_f64_host": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int32), _i, C.POINTER(C.c_double),
                                      C.POINTER(C.c_int32), _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "dafne_resize_bilinear_u8": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "dafne_poly_nms_f64_host": (_i, [C.POINTER(C.c_double), _i, C.c_double, _i, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32)]),
    "dafne_poly_nms_f64_batch_host": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int32), _i, C.c_double, _i,
                                           C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dafne_debug_keep_activations": (_i, [_vp, _i]),
    "dafne_debug_activation": (_i, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "dafne_debug_post_counts": (_i, [_vp, C.POINTER(C.c_int32), _vp]),
    "dafne_debug_nms_stats": (_i, [_vp, C.POINTER(C.c_uint64), _vp]),
    "dafne_set_profiling": (_i, [_vp, _i]),
    "dafne_get_profile": (_i, [_vp, C.c_void_p, _i, C.POINTER(_i)]),
    "dafne_stats": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_double), _i]),
    "dafne_conv_nhwc": (
        _i,
        [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    ),
    "dafne_conv_gn_in_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "dafne_conv1x1_pair_nhwc": (_i, [_vp, C.c_int64, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "dafne_bottleneck_tail_nhwc": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "dafne_gn_relu_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp]),
    "dafne_sort_quadrilateral": (_i, [_vp, _vp, _i, _vp]),
    "dafne_poly_iou": (_i, [_vp, _vp, _vp, _i, _vp]),
    "dafne_poly_pair_filter": (_i, [_vp, _vp, _vp, _i, _vp]),
    "dafne_poly_term_filter": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "dafne_poly_nms": (_i, [_vp, _vp, _vp, _i, _f, _i, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dafne_poly_nms_scratch_bytes": (_i, [_i, C.POINTER(C.c_size_t)]),
    "dafne_poly_nms_host": (_i, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float), _i, _i, _f, _i]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """The loaded shared library; raises DafneError (never falls back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DafneError(
                f"{LIB_PATH} is missing: the CUDA extension is the only execution path. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C dafne_b200/csrc`."
            )
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise DafneError(f"cannot load {LIB_PATH}: {e}") from e
        missing = [name for name in _SIGNATURES if not hasattr(handle, name)]
        if missing:
            raise DafneError(f"{LIB_PATH} does not export {missing}; rebuild it")
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.dafne_abi_version() != 1:
            raise DafneError("libdafne_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().dafne_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str) -> None:
    if status != 0:
        raise DafneError(f"{what} failed: {last_error()}")


def ptr(t) -> int:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(stream=None) -> int:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
