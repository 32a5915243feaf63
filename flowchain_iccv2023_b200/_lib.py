"""ctypes binding of libflowchain_b200.so (C-ABI: include/flowchain_b200.h).

There is NO CPU or PyTorch fallback: if the library is missing or no B200 is visible the
product path raises.  Build the library in-tree with ``python __graft_entry__.py`` (or
``flowchain_iccv2023_b200/csrc/build.sh``)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLOWCHAIN_B200_LIB", os.path.join(_HERE, "libflowchain_b200.so"))

FC_PREC_FP32 = 0
FC_PREC_BF16 = 1
FC_PREC_F16X3 = 2
FC_PREC_F16 = 3
FC_PREFIX_SELF = 0
FC_PREFIX_SHARED = 1

SYMBOLS = [
    "fc_create", "fc_destroy", "fc_last_error", "fc_weight_blob_floats", "fc_load_weights",
    "fc_weights_version", "fc_log_prob", "fc_log_prob_sequential", "fc_grid_log_prob", "fc_grid_log_prob_multicast",
    "fc_sample_with_log_prob", "fc_base_normalizer", "fc_update", "fc_update_lazy", "fc_launch_count", "fc_tc_gemm_selftest",
    "fc_grid_density_maps", "fc_to_trajectories", "fc_best_of_n",
]


class ModelDesc(ctypes.Structure):
    _fields_ = [("pred_len", c_int32), ("input_size", c_int32), ("cond_size", c_int32),
                ("hidden_size", c_int32), ("n_hidden", c_int32), ("n_blocks", c_int32)]


_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the FlowChain B200 CUDA library is not built. Run "
            "`python __graft_entry__.py` (build()) first; there is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    fp = c_void_p  # device pointers are passed as integers
    lib.fc_create.argtypes = [POINTER(c_void_p), POINTER(ModelDesc), c_int]
    lib.fc_create.restype = c_int
    lib.fc_destroy.argtypes = [c_void_p]
    lib.fc_destroy.restype = None
    lib.fc_last_error.argtypes = [c_void_p]
    lib.fc_last_error.restype = c_char_p
    lib.fc_weight_blob_floats.argtypes = [POINTER(ModelDesc)]
    lib.fc_weight_blob_floats.restype = c_size_t
    lib.fc_load_weights.argtypes = [c_void_p, c_void_p, c_size_t, c_uint64, c_void_p]
    lib.fc_load_weights.restype = c_int
    lib.fc_weights_version.argtypes = [c_void_p]
    lib.fc_weights_version.restype = c_uint64
    lib.fc_log_prob.argtypes = [c_void_p, fp, fp, fp, fp, c_uint64, fp, c_int64, c_int, c_void_p]
    lib.fc_log_prob.restype = c_int
    lib.fc_log_prob_sequential.argtypes = [c_void_p, fp, fp, fp, fp, c_uint64, c_int, fp, c_int64, c_int, c_void_p]
    lib.fc_log_prob_sequential.restype = c_int
    lib.fc_grid_log_prob.argtypes = [c_void_p, fp, fp, fp, fp, fp, c_uint64, c_int, c_int, c_int, fp,
                                     c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]
    lib.fc_grid_log_prob.restype = c_int
    lib.fc_grid_log_prob_multicast.argtypes = [c_void_p, fp, fp, fp, fp, fp, c_uint64, c_int, c_int, fp,
                                               c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]
    lib.fc_grid_log_prob_multicast.restype = c_int
    lib.fc_sample_with_log_prob.argtypes = [c_void_p, fp, fp, fp, fp, c_uint64, fp, fp, fp, fp,
                                            c_int64, c_int64, c_int, c_void_p]
    lib.fc_sample_with_log_prob.restype = c_int
    lib.fc_base_normalizer.argtypes = [c_void_p, fp, fp, c_int64, c_int64, c_void_p]
    lib.fc_base_normalizer.restype = c_int
    lib.fc_update.argtypes = [c_void_p, fp, c_int, fp, fp, fp, fp, c_int64, c_int64, c_void_p]
    lib.fc_update.restype = c_int
    lib.fc_update_lazy.argtypes = [c_void_p, fp, c_int, fp, fp, fp, c_int64, c_int64, c_void_p]
    lib.fc_update_lazy.restype = c_int
    lib.fc_grid_density_maps.argtypes = [c_void_p, fp, fp, fp, fp, fp, fp, fp, fp, c_uint64, c_int, fp,
                                         c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]
    lib.fc_grid_density_maps.restype = c_int
    lib.fc_to_trajectories.argtypes = [c_void_p, fp, fp, fp, c_float, c_float, c_int, fp, c_int64, c_int64, c_int, c_int, c_void_p]
    lib.fc_to_trajectories.restype = c_int
    lib.fc_best_of_n.argtypes = [c_void_p, fp, fp, fp, fp, fp, c_float, c_float, c_int, fp, fp, fp, fp, c_int64, c_int64, c_int, c_void_p]
    lib.fc_best_of_n.restype = c_int
    lib.fc_tc_gemm_selftest.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int]
    lib.fc_tc_gemm_selftest.restype = c_int
    lib.fc_launch_count.argtypes = [c_void_p]
    lib.fc_launch_count.restype = c_uint64
    _lib = lib
    return lib


def last_error(handle) -> str:
    msg = load().fc_last_error(handle)
    return msg.decode() if msg else ""
