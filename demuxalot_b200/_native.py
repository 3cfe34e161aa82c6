"""
ctypes binding of libdemux_b200.so (C ABI declared in include/demux_b200.h).

There is no CPU fallback: if the shared library is missing or does not export a declared symbol, importing
the compute path raises.  `python -m demuxalot_b200.build` (or `__graft_entry__.build()`) produces the library.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import LIB_PATH

_i32, _i64, _f32, _f64, _ptr = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p

# symbol -> (restype, argtypes); mirrors include/demux_b200.h one to one
SIGNATURES = {
    'dmx_abi_version': (C.c_int, []),
    'dmx_last_error': (C.c_char_p, []),
    'dmx_device_info': (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(_i64), C.POINTER(_i64)]),
    'dmx_unpack_match_calls': (C.c_int, [_ptr, _i64, _ptr, _i64, _i32, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr]),
    'dmx_host_gather_cb': (C.c_int, [_ptr, _i64, _ptr, _i32]),
    'dmx_build_rows_workspace_bytes': (_i64, [_i64, _i64, _i64]),
    'dmx_build_rows': (C.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _i64,
                                 _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                 C.POINTER(_i64), C.POINTER(_i64), _ptr]),
    'dmx_prior_betas': (C.c_int, [_ptr, _i64, _i64, _i32, _ptr, _ptr, _i64, _ptr, _f64, _ptr, _ptr, _i64, _ptr]),
    'dmx_probs_from_betas': (C.c_int, [_ptr, _i64, _ptr, _i64, _i64, _i32, _ptr, _ptr, _i64, _f32, _f32, _ptr,
                                       _i64, _ptr]),
    'dmx_estep_workspace_bytes': (_i64, [_i64, _i32, _f64, _i64, _i32]),
    'dmx_estep_plan_supported': (C.c_int, [_i32, _f64, _i32]),
    'dmx_estep_plan_workspace_bytes': (_i64, [_i64]),
    'dmx_estep_plan': (C.c_int, [_ptr, _ptr, _i64, _i32, _ptr, _ptr, _i64, _ptr, _i64, C.POINTER(_i64), _ptr]),
    'dmx_barcode_schedule_workspace_bytes': (_i64, [_i64]),
    'dmx_barcode_schedule': (C.c_int, [_ptr, _i64, _ptr, _ptr, _i64, _ptr]),
    'dmx_estep': (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i32, _f64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                            _ptr, _i64, _ptr, _i64, _i32, _f32, _ptr, _ptr, _i64, _i32, _ptr]),
    'dmx_estep_strip_layout': (C.c_int, [_i32, _ptr, C.POINTER(_i32)]),
    'dmx_softmax_rows': (C.c_int, [_ptr, _i64, _i64, _i32, _ptr, _i64, _ptr, _i64, _i32, _ptr]),
    'dmx_mstep': (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i32, _f64, _ptr, _i64, _ptr, _i64, _i64, _i64, _ptr]),
    'dmx_mstep_plan_bytes': (_i64, [_i64]),
    'dmx_mstep_plan': (C.c_int, [_ptr, _i64, _i64, _ptr, _i64, C.POINTER(_i64), _ptr]),
    'dmx_mstep_planned': (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i32, _f64, _ptr, _i64, _ptr, _i64, _i64, _i64,
                                    _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr]),
    'dmx_snp_groups_workspace_bytes': (_i64, [_i64]),
    'dmx_build_snp_groups': (C.c_int, [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr, _i64, _ptr, _ptr, _ptr,
                                       _ptr, C.POINTER(_i64), C.POINTER(_i64), _ptr]),
    'dmx_snp_logits_workspace_bytes': (_i64, [_i64, _i32]),
    'dmx_snp_logits': (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _i32, _f64, _f64, _ptr, _i64, _ptr, _i64,
                                 _ptr]),
    'dmx_softmax_rows_f64': (C.c_int, [_ptr, _i64, _ptr, _i64, _i64, _i32, _ptr, _i64, _ptr, _i64, _i32, _ptr]),
    'dmx_barcode_histogram': (C.c_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr]),
    'dmx_route_calls_workspace_bytes': (_i64, [_i64]),
    'dmx_route_calls': (C.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr, _i32, _ptr, _i64, _ptr, _ptr, _ptr,
                                  C.POINTER(_i64), _ptr]),
    'dmx_peer_sum_f32': (C.c_int, [_ptr, _ptr, _i32, _i32, _i64, _ptr]),
    'dmx_comm_unique_id': (C.c_int, [_ptr]),
    'dmx_comm_init': (C.c_int, [_ptr, _i32, _i32, C.POINTER(_ptr)]),
    'dmx_comm_destroy': (C.c_int, [_ptr]),
    'dmx_mstep_allreduce_padded_variants': (_i64, [_i64, _i32]),
    'dmx_mstep_allreduce': (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i32, _f64, _ptr, _ptr, _ptr, _i64, _ptr, _i64,
                                      _i64, _i64, _i64, _ptr, _ptr, _i32, _i32, _ptr]),
    'dmx_round_f64_to_f32': (C.c_int, [_ptr, _i64, _ptr, _i64, _i64, _i32, _ptr]),
}

ESTEP_EXACT, ESTEP_FAST, ESTEP_AUTO = 0, 1, 2
ABI_VERSION = 4


class NativeError(RuntimeError):
    pass


_lib = None


def load(path: Path = LIB_PATH) -> C.CDLL:
    """Load (once) and type the shared library; raises if it is absent -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not Path(path).exists():
        raise NativeError(
            f'{path} not found: build it with `python -m demuxalot_b200.build` (needs nvcc); '
            'demuxalot_b200 has no CPU fallback')
    lib = C.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing -> loud failure
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.dmx_abi_version() != ABI_VERSION:
        raise NativeError(f'ABI mismatch: library {lib.dmx_abi_version()} vs binding {ABI_VERSION}')
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dmx_last_error().decode(errors='replace')
        raise NativeError(f'{what} failed (rc={rc}): {msg}')


def ptr(tensor) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if tensor is None else tensor.data_ptr()
