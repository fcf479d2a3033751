"""ctypes binding of libb200quant.so (C ABI declared in include/b200quant.h).

Loading never builds silently on a GPU box: the .so is built in-tree by ``__graft_entry__.build()``
(or ``python b200q/build.py``) and shipped as a file.  If it is missing we try one build with nvcc
and otherwise raise — there is no CPU or torch fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

from . import build as _build

_lock = threading.Lock()
_lib = None

c_i64 = C.c_int64
c_vp = C.c_void_p
c_fp = C.c_void_p  # device float* are passed as integers (tensor.data_ptr())

# name -> (restype, argtypes); mirrors include/b200quant.h one to one
SIGNATURES = {
    "b200q_last_error": (C.c_char_p, []),
    "b200q_version": (C.c_int, []),
    "b200q_launch_count": (c_i64, []),
    "b200q_log2_round_threshold_bits": (C.c_uint32, [C.c_int]),
    "b200q_log2_floor_threshold_bits": (C.c_uint32, [C.c_int]),
    "b200q_log2_round_threshold_bits_dt": (C.c_uint32, [C.c_int, C.c_int]),
    "b200q_log2_floor_threshold_bits_dt": (C.c_uint32, [C.c_int, C.c_int]),
    "b200q_col_absmax": (C.c_int, [c_vp, c_i64, c_i64, c_i64, C.c_int, c_fp, C.c_int, c_vp]),
    "b200q_gptq_parity_quant": (C.c_int, [c_vp, c_vp, c_vp, c_fp, c_fp, c_i64, c_i64, c_i64,
                                          C.c_int, C.c_int, c_vp]),
    "b200q_group_fakequant": (C.c_int, [c_vp, c_vp, c_vp, c_fp, c_fp, c_i64, c_i64, c_i64, C.c_int,
                                        C.c_int, C.c_int, c_fp, C.c_int, c_vp]),
    "b200q_smooth_scale": (C.c_int, [c_fp, c_fp, c_fp, c_i64, C.c_float, C.c_int, C.c_int, c_vp]),
    "b200q_col_scale": (C.c_int, [c_vp, c_vp, c_fp, c_i64, c_i64, C.c_int, C.c_int, c_vp]),
    "b200q_act_stat_workspace": (c_i64, [c_i64, c_i64]),
    "b200q_act_meanabs": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_fp, c_vp, c_vp]),
    "b200q_act_meanabs_batched": (C.c_int, [c_vp, C.c_int, c_i64, c_i64, C.c_int, c_fp, c_vp, c_vp]),
    "b200q_act_maxabs": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_fp, C.c_int, c_vp]),
    "b200q_seq_sum_rows": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_fp, c_vp]),
    "b200q_profile_enable": (None, [C.c_int]),
    "b200q_profile_query": (C.c_int, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(c_i64),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200q_smooth_alpha_workspace": (c_i64, [c_i64, c_i64, c_i64, C.c_int]),
    "b200q_smooth_alpha_errors": (C.c_int, [c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, C.c_int, c_vp,
                                            C.c_int, c_vp, c_vp, C.c_int, c_vp]),
    "b200q_packed_words_per_row": (c_i64, [c_i64, C.c_int]),
    "b200q_pack_codes": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp]),
    "b200q_unpack_codes": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp]),
    "b200q_w4a16_gemm": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_vp, c_fp, c_fp, c_i64, c_i64, C.c_int,
                                   c_vp, C.c_int, c_vp]),
    "b200q_selftest_div": (C.c_int, [c_i64, C.c_uint64, C.POINTER(c_i64), c_vp]),
    "b200q_topk_colmul": (C.c_int, [c_fp, c_i64, c_i64, C.c_float, c_fp, c_vp, c_vp]),
    "b200q_awq_layer": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, c_i64, C.c_int,
                                  c_i64, C.c_float, c_fp, c_vp, C.c_int, c_vp]),
    "b200q_gptq_parity_layer": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_fp, C.c_int, c_vp]),
    "b200q_smoothquant_layer": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_int, c_fp, C.c_float,
                                          C.c_int, c_fp, c_fp, C.c_int, c_vp]),
    "b200q_hessian_workspace": (c_i64, [c_i64, c_i64, C.c_int]),
    "b200q_hessian_accum": (C.c_int, [c_vp, C.c_int, c_i64, c_i64, C.c_int, C.c_int, c_fp, C.c_int,
                                      c_fp, c_vp, c_vp]),
    "b200q_awq_search_workspace": (c_i64, [c_i64, c_i64, C.c_int]),
    "b200q_awq_search_loss": (C.c_int, [c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp,
                                        C.POINTER(C.c_float), C.c_int, c_fp, C.c_int, c_vp, c_fp,
                                        c_vp]),
    "b200q_awq_search_loss_folded": (C.c_int, [c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp,
                                               C.POINTER(C.c_float), C.c_int, c_vp, C.c_int, c_vp, c_fp,
                                               c_vp]),
    "b200q_awq_search_delta": (C.c_int, [c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, C.POINTER(C.c_float),
                                         C.c_int, C.c_int, c_vp, c_vp]),
    "b200q_awq_search_loss_prepared": (C.c_int, [c_i64, c_i64, C.c_int, c_fp, c_vp, c_vp, c_fp, c_vp]),
    "b200q_sym_packed_len": (c_i64, [c_i64]),
    "b200q_sym_pack_lower": (C.c_int, [c_fp, c_i64, c_fp, c_vp]),
    "b200q_sym_unpack_lower": (C.c_int, [c_fp, c_i64, c_fp, c_vp]),
    "b200q_sym_fold_packed_bf16": (C.c_int, [c_fp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "b200q_sym_unpack_folded_bf16": (C.c_int, [c_vp, c_i64, c_vp, c_vp]),
    "b200q_hessian_finalize": (C.c_int, [c_fp, c_i64, C.c_float, C.c_float, c_vp]),
    "b200q_spd_inverse_workspace": (c_i64, [c_i64]),
    "b200q_spd_inverse": (C.c_int, [c_fp, c_fp, c_fp, c_i64, c_vp, c_vp, c_vp]),
    "b200q_gptq_compensated_workspace": (c_i64, [c_i64, c_i64]),
    "b200q_gptq_compensated": (C.c_int, [c_fp, c_fp, c_fp, c_i64, c_i64, c_i64, C.c_int, C.c_int,
                                         c_vp, c_vp]),
    "b200q_pot_quant": (C.c_int, [c_vp, c_vp, c_vp, c_fp, c_vp, c_i64, c_i64, C.c_int,
                                  C.POINTER(C.c_float), C.c_int, C.c_int, c_vp]),
    "b200q_apot_quant": (C.c_int, [c_vp, c_vp, c_vp, c_fp, c_vp, c_i64, c_i64,
                                   C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int,
                                   C.c_int, c_vp]),
}


class B200QuantError(RuntimeError):
    """Raised when a libb200quant entry point returns a negative status."""


def lib_path() -> Path:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Return the loaded library, building it in-tree first if the file is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not path.exists():
            _build.build()
        lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the .so disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().b200q_last_error().decode(errors="replace")
        if rc == -1:
            # the reference signals shape errors with assert (quantization_utils.py:384)
            raise AssertionError(f"{what}: {msg}")
        if rc == -3:
            raise NotImplementedError(f"{what}: {msg}")
        raise B200QuantError(f"{what} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(load().b200q_launch_count())


def profile_enable(on: bool) -> None:
    load().b200q_profile_enable(int(on))


def profile_query(name=None):
    """{'ms','launches','bytes','flops'} summed over the recorded calls of entry point `name`."""
    ms, n, by, fl = C.c_double(), c_i64(), C.c_double(), C.c_double()
    rc = load().b200q_profile_query(None if name is None else name.encode(), C.byref(ms),
                                    C.byref(n), C.byref(by), C.byref(fl))
    check(rc, "profile_query")
    return {"ms": ms.value, "launches": n.value, "bytes": by.value, "flops": fl.value}
