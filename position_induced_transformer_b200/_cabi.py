"""ctypes binding of libpit_posatt.so (C ABI declared in include/pit_posatt.h).

Loading never falls back to anything: if the shared library is missing or fails to load the
import raises, and every entry point raises RuntimeError on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libpit_posatt.so")

PIT_EUCLID, PIT_PERIODIC1D, PIT_PERIODIC2D = 0, 1, 2
VARIANT_CODE = {"euclid": PIT_EUCLID, "periodic1d": PIT_PERIODIC1D, "periodic2d": PIT_PERIODIC2D}
ABI_VERSION = 11
PLAN_ROWS, PLAN_COLUMNS = 0, 1

EXPORTS = (
    "pit_abi_version", "pit_last_error", "pit_launch_count", "pit_set_dense_precision", "pit_get_dense_precision", "pit_quantile_ranks", "pit_workspace_bytes",
    "pit_rowstat", "pit_rowstat_lists", "pit_posatt_forward", "pit_posatt_backward", "pit_posatt_backward_coords",
    "pit_decoder_tail_supported", "pit_decoder_tail_forward", "pit_decoder_tail_backward",
    "pit_tail_plan_workspace_bytes", "pit_tail_plan_rows", "pit_tail_plan_fill",
    "pit_head_scale_forward", "pit_head_scale_backward",
    "pit_bias_act_supported", "pit_bias_act_forward", "pit_bias_act_backward",
    "pit_rel_lp_supported", "pit_rel_lp_forward", "pit_rel_lp_backward",
    "pit_processor_supported", "pit_processor_saved_floats", "pit_processor_grad_floats", "pit_processor_scratch_floats",
    "pit_processor_forward", "pit_processor_backward",
    "pit_allreduce_adam_region_floats", "pit_allreduce_adam",
    "pit_mlp_fused_supported", "pit_mlp_fused_forward", "pit_mlp_fused_backward",
)


class Problem(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("variant", "space_dim", "mesh_batched", "batch", "n_head", "n_out", "n_in", "dim")]


class RowStat(C.Structure):
    _fields_ = [("v_min", C.c_void_p), ("v_lo", C.c_void_p), ("v_hi", C.c_void_p),
                ("weight", C.c_float), ("masked", C.c_int32), ("rank_hi", C.c_int32),
                ("nbr_idx", C.c_void_p), ("nbr_d2", C.c_void_p), ("nbr_cnt", C.c_void_p)]


class TailPlan(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("tile_off", C.c_void_p), ("tile_cnt", C.c_void_p), ("cand", C.c_void_p), ("d2", C.c_void_p),
                ("n_tiles", C.c_int32)]


class ProcessorBlock(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p)]


class AllReduceAdam(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("n_tensors", C.c_int32), ("grad", C.c_void_p * 64), ("numel", C.c_int32 * 64),
                ("region", C.c_void_p * 16), ("param", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("step", C.c_void_p), ("sync", C.c_void_p), ("lr", C.c_void_p), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float)]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m position_induced_transformer_b200.build` "
            "(there is no CPU or PyTorch fallback for position-attention)")
    lib = C.CDLL(LIB_PATH)
    p, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    lib.pit_abi_version.restype = C.c_int
    lib.pit_set_dense_precision.argtypes = [i32]
    lib.pit_last_error.restype = C.c_char_p
    lib.pit_launch_count.restype = C.c_uint64
    lib.pit_quantile_ranks.argtypes = [C.c_double, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_float)]
    lib.pit_workspace_bytes.argtypes = [C.POINTER(Problem)]
    lib.pit_workspace_bytes.restype = C.c_size_t
    lib.pit_rowstat.argtypes = [C.POINTER(Problem), f32p, f32p, f32p, i32, i32, f32p, f32p, f32p, p]
    lib.pit_rowstat_lists.argtypes = [C.POINTER(Problem), f32p, f32p, f32p, i32, i32, f32p, f32p, f32p, p, p, p, p]
    lib.pit_posatt_forward.argtypes = [C.POINTER(Problem), f32p, f32p, f32p, f32p, f32p, C.POINTER(RowStat),
                                       f32p, i64, i64, i32, f32p, p, C.c_size_t, C.POINTER(TailPlan), p]
    lib.pit_posatt_backward.argtypes = [C.POINTER(Problem), f32p, f32p, f32p, f32p, f32p, C.POINTER(RowStat), f32p,
                                        f32p, i64, i64, i32, f32p, f32p, p, C.c_size_t, C.POINTER(TailPlan), p]
    lib.pit_posatt_backward_coords.argtypes = [C.POINTER(Problem), f32p, f32p, f32p, f32p, f32p, C.POINTER(RowStat), f32p,
                                               f32p, i64, i64, f32p, f32p, f32p, p]
    lib.pit_decoder_tail_supported.argtypes = [C.POINTER(Problem), i32]
    lib.pit_decoder_tail_forward.argtypes = [C.POINTER(Problem), f32p, f32p, f32p, f32p, f32p, C.POINTER(RowStat),
                                             f32p, f32p, f32p, i32, f32p, f32p, C.POINTER(TailPlan), p]
    lib.pit_decoder_tail_backward.argtypes = [C.POINTER(Problem), f32p, f32p, f32p, f32p, f32p, C.POINTER(RowStat),
                                              f32p, f32p, f32p, i32, f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                              C.POINTER(TailPlan), p]
    lib.pit_tail_plan_workspace_bytes.argtypes = [C.POINTER(Problem), i32]
    lib.pit_tail_plan_workspace_bytes.restype = C.c_size_t
    lib.pit_tail_plan_rows.argtypes = [C.POINTER(Problem), i32, f32p, f32p, f32p, C.POINTER(RowStat), p, p, p, C.c_size_t, p]
    lib.pit_tail_plan_fill.argtypes = [C.POINTER(Problem), i32, f32p, f32p, f32p, C.POINTER(RowStat), p, p, p, p, p, C.c_size_t, p]
    lib.pit_head_scale_forward.argtypes = [f32p, f32p, i32, p]
    lib.pit_head_scale_backward.argtypes = [f32p, f32p, f32p, f32p, i32, p]
    lib.pit_bias_act_supported.argtypes = [i64, i32]
    lib.pit_bias_act_forward.argtypes = [f32p, f32p, f32p, i64, i32, i32, p]
    lib.pit_bias_act_backward.argtypes = [f32p, f32p, f32p, f32p, f32p, i64, i32, i32, p]
    lib.pit_rel_lp_supported.argtypes = [i32, i64, i32, i32]
    lib.pit_rel_lp_forward.argtypes = [f32p, f32p, i32, i64, i32, i32, f32p, f32p, p]
    lib.pit_rel_lp_backward.argtypes = [f32p, f32p, f32p, f32p, i32, i64, i32, i32, f32p, p]
    lib.pit_processor_supported.argtypes = [C.POINTER(Problem), i32]
    for name in ("pit_processor_saved_floats", "pit_processor_grad_floats"):
        getattr(lib, name).argtypes = [C.POINTER(Problem), i32]
        getattr(lib, name).restype = C.c_size_t
    lib.pit_processor_scratch_floats.argtypes = [C.POINTER(Problem)]
    lib.pit_processor_scratch_floats.restype = C.c_size_t
    lib.pit_processor_forward.argtypes = [C.POINTER(Problem), i32, f32p, f32p, f32p, f32p, C.POINTER(ProcessorBlock), i32, f32p, f32p, p]
    lib.pit_processor_backward.argtypes = [C.POINTER(Problem), i32, f32p, f32p, f32p, f32p, C.POINTER(ProcessorBlock), i32, f32p,
                                           f32p, f32p, f32p, f32p, p]
    lib.pit_allreduce_adam_region_floats.argtypes = [i64]
    lib.pit_allreduce_adam_region_floats.restype = C.c_size_t
    lib.pit_allreduce_adam.argtypes = [C.POINTER(AllReduceAdam), p]
    lib.pit_mlp_fused_supported.argtypes = [i64, i32, i32, i32]
    lib.pit_mlp_fused_forward.argtypes = [f32p, f32p, f32p, f32p, f32p, i64, i32, i32, i32, i32, f32p, f32p, f32p, p]
    lib.pit_mlp_fused_backward.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, i64, i32, i32, i32, i32, f32p, f32p, p]
    if lib.pit_abi_version() != ABI_VERSION:
        raise ImportError(f"libpit_posatt.so ABI {lib.pit_abi_version()} != expected {ABI_VERSION}; rebuild it")
    return lib


lib = _load()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib.pit_last_error().decode()}")


def quantile_ranks(q: float, m: int):
    """(k_lo, k_hi, w) of torch.quantile(x, q, dim=-1) on rows of m entries."""
    lo, hi, w = C.c_int32(), C.c_int32(), C.c_float()
    check(lib.pit_quantile_ranks(float(q), int(m), C.byref(lo), C.byref(hi), C.byref(w)), "pit_quantile_ranks")
    return lo.value, hi.value, w.value


def launch_count() -> int:
    return int(lib.pit_launch_count())
