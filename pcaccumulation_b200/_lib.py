"""ctypes binding of libpcab200.so (the C ABI declared in include/pcab200.h).

The product path has NO fallback: if the shared library is missing or a call fails, an exception is
raised.  Only raw device pointers, sizes and the current CUDA stream cross this boundary.
"""
import ctypes
import threading
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcab200.so")

_lib = None


class PcabError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcabError(
                f"{LIB_PATH} is missing: build it with `python -m pcaccumulation_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pcab_last_error.restype = ctypes.c_char_p
        for name in set(SIZE_T_FUNCS) | {n for n in SYMBOLS if n.endswith("_workspace")}:
            getattr(_lib, name).restype = ctypes.c_size_t
    return _lib


SIZE_T_FUNCS = [
    "pcab_voxelize_workspace", "pcab_pillar_index_workspace", "pcab_pillar_encode_workspace",
    "pcab_bg_compact_workspace", "pcab_ego_pairs_workspace", "pcab_select_workspace", "pcab_cluster_workspace",
    "pcab_tpn_iteration_workspace", "pcab_chamfer_workspace", "pcab_conv3x3_tc_pack_floats", "pcab_tpn_rows_workspace",
    "pcab_stpn_head_tc_pack_floats", "pcab_prep_points_workspace", "pcab_cluster_eval_workspace",
]

# every symbol include/pcab200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "pcab_last_error", "pcab_version", "pcab_voxelize_workspace", "pcab_voxelize", "pcab_pillar_index_workspace",
    "pcab_pillar_index", "pcab_pillar_stats", "pcab_pillar_cells", "pcab_pfn_pack_size",
    "pcab_pillar_encode_workspace", "pcab_pillar_encode", "pcab_pillar_encode_tc", "pcab_conv3x3_f32", "pcab_convT2x2_f32", "pcab_maxpool2x2",
    "pcab_temporal_max", "pcab_conv3x3_tc_supported", "pcab_conv3x3_tc_plan", "pcab_conv3x3_tc_pack_floats", "pcab_conv3x3_tc", "pcab_conv3x3_tc_f16",
    "pcab_conv3x3_p16_supported", "pcab_conv_p16_plan", "pcab_conv3x3_p16", "pcab_conv3d_p16", "pcab_convT2x2_p16",
    "pcab_head2_conv", "pcab_fb_per_point", "pcab_canvases", "pcab_warp_bev", "pcab_transform_points",
    "pcab_bg_compact_workspace", "pcab_bg_compact", "pcab_ego_pairs_workspace", "pcab_ego_pairs",
    "pcab_select_workspace", "pcab_select_indices", "pcab_ungrid", "pcab_stpn_head_pack_size",
    "pcab_init_point_outputs", "pcab_stpn_head", "pcab_stpn_head_tc_pack_floats", "pcab_stpn_head_tc", "pcab_dynamic_flags", "pcab_cluster_workspace", "pcab_cluster_scene",
    "pcab_tpn_relabel", "pcab_tpn_rows_workspace", "pcab_tpn_rows", "pcab_tpn_static_embed", "pcab_tpn_iteration_workspace", "pcab_tpn_iteration", "pcab_embed_segmax_tc", "pcab_tpn_pos_l0", "pcab_apply_seg_pose",
    "pcab_scatter_rows3", "pcab_tpn_gather", "pcab_pose_error", "pcab_inst_errors", "pcab_prep_points_workspace", "pcab_prep_points", "pcab_prep_points_augmented", "pcab_flow_eval", "pcab_cluster_eval_workspace", "pcab_cluster_eval", "pcab_chamfer_workspace", "pcab_chamfer_forward", "pcab_chamfer_backward",
    "pcab_nn_workspace", "pcab_nn_search", "pcab_icp_workspace", "pcab_icp_point_to_point", "pcab_ego_pose_errors", "pcab_chamfer_forward_brute",
    "pcab_seg_loss_workspace", "pcab_seg_loss", "pcab_seg_loss_grad", "pcab_offset_loss_workspace", "pcab_offset_loss",
    "pcab_offset_loss_grad", "pcab_perm_loss",
]


def P(t):
    """Device (or host, for numpy-backed ctypes arrays) pointer argument."""
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.cast(t, ctypes.c_void_p)


I = ctypes.c_int
F = ctypes.c_float
D = ctypes.c_double
L = ctypes.c_longlong
Z = ctypes.c_size_t


def host_floats(values):
    return (ctypes.c_float * len(values))(*[float(v) for v in values])


_tls = threading.local()


def stream():
    """Raw handle of torch's current stream.  Inside ``pinned_stream()`` the handle fetched at its entry is reused
    (``torch.cuda.current_stream()`` costs ~15 us and a forward makes ~100 calls)."""
    s = getattr(_tls, "stream", None)
    if s is not None:
        return s
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class pinned_stream:
    """Context manager: every ``stream()`` call of this thread returns the stream that is current at entry."""

    def __enter__(self):
        self.prev = getattr(_tls, "stream", None)
        _tls.stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        return self

    def __exit__(self, *exc):
        _tls.stream = self.prev
        return False


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise PcabError(f"{name} failed ({rc}): {lib().pcab_last_error().decode()}")


def size(name, *args):
    return int(getattr(lib(), name)(*args))


def scratch(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
