"""ctypes binding of include/stemgnn_b200.h.  The library is mandatory: there is no Python/CPU
fallback for the hot path — if the shared object is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libstemgnn_b200.so")

MAX_STACK = 2
ABI_VERSION = 3


class Dims(Structure):
    _fields_ = [("B", c_int), ("N", c_int), ("W", c_int), ("H", c_int), ("multi", c_int)]


class BlockPtrs(Structure):
    """stemgnn_block_params_t / stemgnn_block_grads_t (identical layout: all pointers)."""
    _fields_ = [("weight", c_void_p), ("forecast_w", c_void_p), ("forecast_b", c_void_p),
                ("forecast_result_w", c_void_p), ("forecast_result_b", c_void_p),
                ("backcast_w", c_void_p), ("backcast_b", c_void_p), ("shortcut_w", c_void_p),
                ("shortcut_b", c_void_p), ("glu_left_w", c_void_p * 6), ("glu_left_b", c_void_p * 6),
                ("glu_right_w", c_void_p * 6), ("glu_right_b", c_void_p * 6)]


class ModelPtrs(Structure):
    """stemgnn_params_t / stemgnn_grads_t."""
    _fields_ = [("weight_key", c_void_p), ("weight_query", c_void_p), ("gru_w_ih", c_void_p),
                ("gru_w_hh", c_void_p), ("gru_b_ih", c_void_p), ("gru_b_hh", c_void_p),
                ("block", BlockPtrs * MAX_STACK), ("fc0_w", c_void_p), ("fc0_b", c_void_p),
                ("fc2_w", c_void_p), ("fc2_b", c_void_p)]


# stemgnn_allreduce_fn(dev_buf, n, user, stream)
ALLREDUCE_FN = ctypes.CFUNCTYPE(None, c_void_p, ctypes.c_longlong, c_void_p, c_void_p)


class FwdOpts(Structure):
    _fields_ = [("leaky_alpha", c_float), ("dropout_p", c_float), ("training", c_int),
                ("dropout_seed", c_uint64), ("dropout_offset", c_uint64), ("dropout_mask", c_void_p),
                ("gemm_mode", c_int), ("reuse_folded", c_int), ("graph_mode", c_int), ("dropout_offset_dev", c_void_p),
                ("graph_allreduce", ALLREDUCE_FN), ("graph_allreduce_user", c_void_p)]


# name -> (restype, argtypes); every symbol include/stemgnn_b200.h declares
SYMBOLS = {
    "stemgnn_version": (c_int, []),
    "stemgnn_last_error": (c_char_p, []),
    "stemgnn_device_ok": (c_int, []),
    "stemgnn_launch_count": (ctypes.c_longlong, []),
    "stemgnn_profile_gru": (None, [c_void_p, c_void_p]),
    "stemgnn_gru_kernel_name": (c_char_p, []),
    "stemgnn_workspace_bytes": (c_size_t, [POINTER(Dims), c_int]),
    "stemgnn_model_forward": (c_int, [POINTER(Dims), POINTER(ModelPtrs), POINTER(FwdOpts), c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemgnn_model_backward": (c_int, [POINTER(Dims), POINTER(ModelPtrs), POINTER(FwdOpts), c_void_p,
                                       c_void_p, c_void_p, POINTER(ModelPtrs), c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    "stemgnn_gru_keyquery_forward": (c_int, [POINTER(Dims), POINTER(ModelPtrs), c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "stemgnn_graph_forward": (c_int, [POINTER(Dims), POINTER(FwdOpts), c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemgnn_block_forward": (c_int, [POINTER(Dims), POINTER(BlockPtrs), c_int, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemgnn_spe_seq_cell_forward": (c_int, [POINTER(Dims), POINTER(BlockPtrs), c_int, c_void_p,
                                             c_void_p, c_void_p, c_size_t, c_void_p]),
    "stemgnn_gather_windows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p]),
    "stemgnn_eval_metrics": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_int, c_void_p, c_void_p]),
    "stemgnn_laplacian_eig_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float,
                                              c_void_p]),
    "stemgnn_glu_chain": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                  c_void_p]),
    "stemgnn_glu_chain_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "stemgnn_mse_loss_grad": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_void_p]),
    "stemgnn_optimizer_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_float,
                                       c_float, c_float, c_void_p, c_void_p]),
    "stemgnn_counters_tick": (c_int, [c_void_p, c_void_p, ctypes.c_ulonglong, c_void_p]),
    "stemgnn_sgemm": (c_int, [c_int, c_int, c_int, c_float, c_void_p, c_int, c_int, c_void_p, c_int,
                              c_int, c_float, c_void_p, c_int, c_void_p]),
    "stemgnn_tc_gemm": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "stemgnn_gft_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "stemgnn_glu_gemm": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_void_p]),
}

_lib = None


def load():
    """Loads libstemgnn_b200.so (once).  Raises RuntimeError when it is absent — build it with
    `python -m stemgnn_b200.build` (or `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"stemgnn_b200: {LIB_PATH} not found. The CUDA library is required (no CPU fallback); "
            "build it with `python -m stemgnn_b200.build`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError here = ABI mismatch, let it propagate
        fn.restype = res
        fn.argtypes = args
    if lib.stemgnn_version() != ABI_VERSION:
        raise RuntimeError(f"stemgnn_b200: ABI version {lib.stemgnn_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().stemgnn_last_error().decode(errors="replace")
        raise RuntimeError(f"stemgnn_b200: {what} failed (rc={rc}): {msg}")
