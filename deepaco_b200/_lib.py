"""ctypes binding of libdeepaco_b200.so (the C ABI declared in include/deepaco_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is
raised.  torch is used for device memory, streams and the Philox generator state only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeepaco_b200.so")

_u64, _i64, _i32, _f32, _vp = C.c_uint64, C.c_int64, C.c_int, C.c_float, C.c_void_p

_SIGNATURES = {
    # name: (restype, [argtypes])
    "deepaco_last_error": (C.c_char_p, []),
    "deepaco_version": (_i32, []),
    "deepaco_kernel_launches": (C.c_longlong, []),
    "deepaco_torch_draw_geometry": (_i32, [_i64, C.POINTER(C.c_uint32), C.POINTER(_u64)]),
    "deepaco_aten_sum_plan": (_i32, [_i32, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "deepaco_tsp_sample": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "deepaco_tsp_sample_shard": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]),
    "deepaco_tsp_sample_shard_p2p": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _u64, _u64, _vp, _vp, _i32, _i32, _vp, _i32, _vp]),
    "deepaco_tsp_sample_offset_increment": (_u64, [_i32, _i32, _i32]),
    "deepaco_tsp_roulette_sample": (_i32, [_vp, _i32, _i32, _i32, _i32, _u64, _u64, _vp, _vp, _vp, _vp]),
    "deepaco_tsp_roulette_offset_increment": (_u64, [_i32, _i32]),
    "deepaco_tsp_cost": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "deepaco_tsp_update": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _f32, _vp, _vp]),
    "deepaco_tsp_update_tours": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _f32, _vp, _vp]),
    "deepaco_two_opt": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "deepaco_tsp_nls": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "deepaco_paths_to_tours": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "deepaco_tours_to_paths": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "deepaco_knn_graph": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "deepaco_gnn_weight_count": (_i64, [_i32]),
    "deepaco_gnn_forward": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _vp, _vp]),
    "deepaco_cvrp_sample": (_i32, [_vp, _vp, _vp, _f32, _i32, _i32, _i32, _u64, _u64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "deepaco_cvrp_step_offset_increment": (_u64, [_i32, _i32]),
    "deepaco_cvrp_cost": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "deepaco_cvrp_update": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32, _i32, _f32, _vp, _vp]),
    "deepaco_pick_move": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _u64, _u64, _vp, _vp, _vp, _vp]),
    "deepaco_pick_move_offset_increment": (_u64, [_i32, _i32]),
    "deepaco_logp_backward": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _f32, _vp, _vp, _vp]),
    "deepaco_debug_exponential": (_i32, [_u64, _u64, _i64, _vp, _vp]),
    "deepaco_debug_randint": (_i32, [_u64, _u64, _i64, _i64, _vp, _vp]),
    "deepaco_debug_row_sum": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "deepaco_debug_exp_guard": (_i32, [_vp, _vp]),
}



class TspRunArgs(C.Structure):
    """deepaco_tsp_run_args (include/deepaco_b200.h)."""
    _fields_ = [("n", _i32), ("n_ants", _i32), ("n_colonies", _i32), ("start_node", _i32), ("double_norm", _i32),
                ("decay", _f32), ("elitist", _i32), ("min_max", _i32), ("ph_min", _f32),
                ("seed", _u64), ("offset", _u64), ("offsets", _vp),
                ("pheromone", _vp), ("heuristic", _vp), ("distances", _vp), ("product", _vp), ("product_valid", _i32),
                ("tours", _vp), ("costs", _vp), ("neighbours", _vp), ("lowest_cost", _vp), ("shortest_path", _vp),
                ("ph_max", _vp), ("scale", _vp), ("knn", _vp), ("local_search", _i32), ("ls_max_iterations", _i32),
                ("T_nls", _i32), ("T_p", _i32), ("heuristic_dist", _vp), ("ev_sample_begin", _vp), ("ev_sample_end", _vp),
                ("knn_refresh", _i32), ("knn_iteration0", _i32), ("roulette", _i32)]


class ShardArgs(C.Structure):
    """deepaco_shard_args (include/deepaco_b200.h)."""
    _fields_ = [("rank", _i32), ("world", _i32), ("ant_base", _i32), ("n_ants_local", _i32), ("peer_tours_host", _vp),
                ("peer_flags_host", _vp), ("epoch", C.c_uint32), ("timeout_ms", C.c_uint32), ("status", _vp)]


class CvrpRunArgs(C.Structure):
    """deepaco_cvrp_run_args (include/deepaco_b200.h)."""
    _fields_ = [("n_nodes", _i32), ("n_ants", _i32), ("n_colonies", _i32), ("capacity", _f32), ("decay", _f32),
                ("elitist", _i32), ("min_max", _i32), ("ph_min", _f32), ("seed", _u64), ("offsets", _vp),
                ("pheromone", _vp), ("heuristic", _vp), ("distances", _vp), ("demand", _vp), ("product", _vp),
                ("product_valid", _i32), ("tours", _vp), ("costs", _vp), ("neighbours", _vp), ("lens", _vp), ("tmax", _vp),
                ("lowest_cost", _vp), ("shortest_path", _vp), ("shortest_rows", _vp), ("ph_max", _vp), ("scale", _vp)]


class GnnTrainArgs(C.Structure):
    """deepaco_gnn_train_args (include/deepaco_b200.h)."""
    _fields_ = [("n_nodes", _i32), ("n_edges", _i32), ("feats", _i32), ("n_instances", _i32), ("ctas_per_instance", _i32),
                ("bn_eps", _f32), ("x", _vp), ("row_ptr", _vp), ("src_sorted", _vp), ("dst_sorted", _vp), ("attr_sorted", _vp),
                ("order", _vp), ("col_ptr", _vp), ("in_edges", _vp), ("weights", _vp), ("xs", _vp), ("ws", _vp), ("zv", _vp),
                ("ze", _vp), ("stats", _vp), ("node_ws", _vp), ("edge_ws", _vp), ("red", _vp), ("sync_ws", _vp), ("heu_out", _vp),
                ("grad_heu", _vp), ("grad_weights", _vp)]


_SIGNATURES["deepaco_gnn_train_forward"] = (_i32, [C.POINTER(GnnTrainArgs), _vp])
_SIGNATURES["deepaco_gnn_train_backward"] = (_i32, [C.POINTER(GnnTrainArgs), _vp])
_SIGNATURES["deepaco_gnn_forward_group"] = (_i32, [C.POINTER(GnnTrainArgs), _vp])
_SIGNATURES["deepaco_cvrp_run"] = (_i32, [C.POINTER(CvrpRunArgs), _i32, _vp])
_SIGNATURES["deepaco_tsp_run"] = (_i32, [C.POINTER(TspRunArgs), _i32, _vp])
_SIGNATURES["deepaco_tsp_run_shard"] = (_i32, [C.POINTER(TspRunArgs), C.POINTER(ShardArgs), _i32, _vp])
_SIGNATURES["deepaco_tsp_run_host"] = (_i32, [C.POINTER(TspRunArgs), _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp])

_lib = None


class DeepAcoError(RuntimeError):
    pass


def exported_symbols():
    """Names the header declares; tests check each resolves in the built library."""
    return list(_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DeepAcoError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C deepaco_b200/csrc`). deepaco_b200 has no CPU fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().deepaco_last_error().decode(errors="replace")
        raise DeepAcoError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise DeepAcoError(f"deepaco_b200: `{name}` must live on a CUDA device (got {t.device}); "
                           "this engine has no CPU path")
    return t


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous view/copy (no-op for the tensors the reference drivers pass)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def generator_state(device):
    """(seed, offset) of torch's default CUDA generator on `device`."""
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    g = torch.cuda.default_generators[idx]
    return g, int(g.initial_seed()), int(g.get_offset())
