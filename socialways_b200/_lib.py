"""ctypes binding of the C-ABI library (include/socialways_b200.h).

There is NO fallback: if libsocialways_b200.so is missing or a call fails, this raises.  Build the
library with `python -m socialways_b200.build` (or __graft_entry__.build()).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SOCIALWAYS_B200_LIB: an alternative build of the SAME library (A/B experiments with compile-time constants)
LIB_PATH = os.environ.get("SOCIALWAYS_B200_LIB") or os.path.join(HERE, "libsocialways_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_D = ctypes.c_double

_PROTOTYPES = {
    "sw_abi_version": (_I, []),
    "sw_last_cuda_error": (_I, []),
    "sw_error_string": (ctypes.c_char_p, [_I]),
    "sw_decode_pack_floats": (_I, []),
    "sw_pool_pack_floats": (_I, []),
    "sw_decode_pack_t_floats": (_I, []),
    "sw_lstm_seq_fwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "sw_lstm_seq_fwd_tcx": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "sw_lstm_seq_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "sw_pool_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "sw_pool_bwd": (_I, [_P] * 17 + [_I, _I, _P]),
    "sw_pool_fwd_tcx": (_I, [_P] * 9 + [_I, _I, _I, _I, _I, _P]),
    "sw_pool_tcx_max_scene": (_I, []),
    "sw_decode_fwd": (_I, [_P] * 13 + [_I, _I, _I, _I, _P]),
    "sw_decode_bwd": (_I, [_P] * 16 + [_I, _I, _I, _I, _P]),
    "sw_decode_fwd_tc": (_I, [_P] * 8 + [_I, _I, _I, _I, _P]),
    "sw_decode_tc_pack_sizes": (_I, [_P, _P]),
    "sw_decode_fwd_tcx": (_I, [_P] * 10 + [_I, _I, _I, _I, _P]),
    "sw_decode_tcx_pack_sizes": (_I, [_P, _P, _P]),
    "sw_decode_fwd_pair": (_I, [_P] * 9 + [ctypes.c_longlong, _P, _I, _I, _I, _I, _P]),
    "sw_decode_fwd_pair_bf16": (_I, [_P] * 9 + [ctypes.c_longlong, _P, _I, _I, _I, _I, _P]),
    "sw_decode_pair_pack_sizes": (_I, [_P, _P]),
    "sw_decode_pair_scratch_bytes": (ctypes.c_longlong, [_I]),
    "sw_disc_heads_pack_floats": (_I, [_I, _I]),
    "sw_disc_heads_record_dims": (_I, [_I, _I, _P, _P]),
    "sw_disc_heads_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "sw_disc_heads_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "sw_gen_pack_sizes": (_I, [_P] * 7),
    "sw_gen_pack": (_I, [_P] * 9),
    "sw_gen_pack_bwd": (_I, [_P] * 5 + [_I, _P]),
    "sw_disc_pack_sizes": (_I, [_I, _P, _P, _P]),
    "sw_disc_pack": (_I, [_P, _I, _P, _P, _P, _P]),
    "sw_disc_step_image_rows": (_I, [_I, _P, _P]),
    "sw_disc_step": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _P, _I, _P, _F, _F] + [_P] * 7 + [_I, _I, _P]),
    "sw_contract_plan": (_I, [_P, _I, _I, _I, _P, _P]),
    "sw_contract": (_I, [_P, _I, _P, ctypes.c_longlong, _P, _I, _I, _P]),
    "sw_contract_tc": (_I, [_P, _I, _P, ctypes.c_longlong, _P, _I, _I, _P]),
    "sw_rows_linear": (_I, [_P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "sw_train_stats": (_I, [_P, _P, _I, _I, _F, _P, _I, _P, _I, _F, _F, _P, _P, _P, _I, _P]),
    "sw_noise_uniform": (_I, [_P, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_ulonglong, _I, _P]),
    "sw_bestofk_metrics": (_I, [_P, _P, _F, _I, _I, _I, _P, _P]),
    "sw_traj_nn1_counts": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "sw_traj_emd_cost": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "sw_lsap_smem_bytes": (_I, [_I]),
    "sw_lsap_solve": (_I, [_P, _I, _I, _P, _P, _P]),
    "sw_adam_flat": (_I, [_P, _P, _P, _P, _P, _I, _D, _D, _D, _D, _I, _P]),
    "sw_set_peer_wait_timeout_ms": (_I, [_I]),
    "sw_allreduce_adam": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _D, _D, _D, _D, _P]),
}

_lib = None


class SocialWaysCudaError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SocialWaysCudaError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m socialways_b200.build`.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def check(code, what):
    if code != 0:
        msg = lib().sw_error_string(code).decode()
        raise SocialWaysCudaError(f"{what} failed: {msg} (code {code})")


def ptr(t):
    """Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SocialWaysCudaError("socialways_b200 kernels take CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise SocialWaysCudaError("tensor must be contiguous")
    return t.data_ptr()
