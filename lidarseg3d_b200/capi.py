"""ctypes binding of the C-ABI shared library (include/ls3d.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module without a
built ``_ls3d.so`` raises, and every wrapper raises ``RuntimeError`` on a non-zero status.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ls3d.so")
if os.environ.get("LS3D_PROF_SO") == "1":       # development: instrumented build (build.build(prof=True))
    LIB_PATH = os.path.join(_HERE, "_ls3d_prof.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int32
c_float = ctypes.c_float


class GemmArgs(ctypes.Structure):
    """Mirror of ``ls3d_gemm_args`` (include/ls3d.h)."""

    _fields_ = [
        ("in0", c_void_p), ("in1", c_void_p),
        ("ld0", c_int), ("c0", c_int), ("ld1", c_int), ("c1", c_int),
        ("nbr", c_void_p),
        ("koff", c_int), ("m_out", c_int),
        ("w", c_void_p),
        ("cin_pad", c_int), ("n_pad", c_int), ("cout", c_int), ("epi", c_int),
        ("scale", c_void_p), ("shift", c_void_p),
        ("relu", c_int),
        ("res", c_void_p), ("ld_res", c_int), ("res_mode", c_int),
        ("red0", c_void_p), ("red1", c_void_p),
        ("ld_red0", c_int), ("ld_red1", c_int), ("red_c", c_int),
        ("n_ln", c_int),
        ("ln_g0", c_void_p), ("ln_b0", c_void_p), ("ln_g1", c_void_p), ("ln_b1", c_void_p),
        ("ln_eps", c_float),
        ("attn_k", c_void_p), ("attn_v", c_void_p), ("frame_off", c_void_p),
        ("n_frames", c_int), ("n_tok", c_int), ("n_head", c_int),
        ("attn_scale", c_float),
        ("row_mask", c_void_p), ("ld_mask", c_int),
        ("out", c_void_p), ("ld_out", c_int), ("round_out", c_int), ("debug_skip", c_int), ("precise", c_int),
        ("plan_hdr", c_void_p), ("plan_local", c_void_p), ("plan_pool", c_void_p),
    ]


_lib = None


def lib():
    """Load (once) and return the shared library; fail loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(lidarseg3d_b200 has no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.ls3d_gather_gemm.argtypes = [ctypes.POINTER(GemmArgs), c_void_p]
    L.ls3d_gather_gemm.restype = ctypes.c_int
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = restype


class ConvArgs(ctypes.Structure):
    """ls3d_conv_args (include/ls3d.h)."""
    _fields_ = [("in16", c_void_p), ("w_packed", c_void_p), ("bias", c_void_p), ("res32", c_void_p), ("out32", c_void_p),
                ("out16", c_void_p), ("in_c_total", c_int), ("in_c_off", c_int), ("cin", c_int), ("out_c_total", c_int),
                ("out_c_off", c_int), ("cout", c_int), ("n_img", c_int), ("H_in", c_int), ("W_in", c_int), ("ksize", c_int),
                ("stride", c_int), ("relu", c_int), ("w_split", c_int), ("res16", c_void_p)]


class ConvPass(ctypes.Structure):
    """ls3d_conv_pass (include/ls3d.h)."""
    _fields_ = [("w_packed", c_void_p), ("in_c_off", c_int), ("out_c_off", c_int), ("flags", c_int)]


P, I, L = c_void_p, c_int, ctypes.c_int64
PL = ctypes.POINTER(ctypes.c_int64)
# name -> (argtypes, restype): one row per prototype of include/ls3d.h
_SIGNATURES = {
    "ls3d_gemm_packed_bytes": ([I, I, I, PL, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)], ctypes.c_int),
    "ls3d_gemm_pack_bf16x3": ([P, I, I, I, P, P], ctypes.c_int),
    "ls3d_tile_plan_bytes": ([I, I, PL, PL, PL], ctypes.c_int),
    "ls3d_tile_plan_build": ([P, I, I, P, P, P, P, P], ctypes.c_int),
    "ls3d_voxelize_workspace_bytes": ([L, I, I, PL], ctypes.c_int),
    "ls3d_voxelize": ([P, I, I, P, I, P, P, I, I, P, L, P, P, P, P, P, P, P], ctypes.c_int),
    "ls3d_vfe_descriptor": ([P, P, I, I, I, I, P, I, I, P], ctypes.c_int),
    "ls3d_vfe_token_attn": ([P, I, I, I, I, I, P, I, I, P], ctypes.c_int),
    "ls3d_vfe_token_max": ([P, I, I, I, I, P, I, I, P], ctypes.c_int),
    "ls3d_grid_bytes": ([I, I, I, I, PL, PL], ctypes.c_int),
    "ls3d_grid_build": ([P, I, I, I, I, I, P, P, P, P, P], ctypes.c_int),
    "ls3d_grid_build_strided": ([P, I, I, P, P, P, I, I, I, P, P, P, P], ctypes.c_int),
    "ls3d_grid_enumerate": ([P, I, I, I, I, P, P], ctypes.c_int),
    "ls3d_rulebook_gather": ([P, P, I, I, I, I, P, I, P, P, P, P, P], ctypes.c_int),
    "ls3d_rulebook_scatter": ([P, I, I, I, I, P, I, P, P, P, P, P], ctypes.c_int),
    "ls3d_three_nn_grid": ([P, I, I, P, P, I, I, I, I, P, P, P, P, P, P, P, P, P, P], ctypes.c_int),
    "ls3d_three_nn": ([I, I, I, P, P, P, P, P], ctypes.c_int),
    "ls3d_frame_offsets": ([P, I, L, I, I, P, P], ctypes.c_int),
    "ls3d_three_interpolate": ([P, I, I, P, P, I, P, I, I, P], ctypes.c_int),
    "ls3d_sample_image_features": ([P, I, I, I, I, I, I, P, I, P, P, I, I, P], ctypes.c_int),
    "ls3d_project_points": ([P, I, I, I, P, P, I, I, I, I, I, P, P], ctypes.c_int),
    "ls3d_project_points_global": ([P, I, I, I, P, P, P, I, I, I, I, I, P, P], ctypes.c_int),
    "ls3d_resize_images_u8": ([P, I, I, I, I, I, P, P, P, I, P], ctypes.c_int),
    "ls3d_sffm_decoder_weight_bytes": ([I, PL, PL], ctypes.c_int),
    "ls3d_sffm_decoder": ([P, I, I, P, P, P, P, P, I, I, I, I, I, I, I, ctypes.c_float, ctypes.c_float, P, I, P], ctypes.c_int),
    "ls3d_token_attention": ([P, I, I, P, P, P, I, I, I, I, ctypes.c_float, P, I, I, P], ctypes.c_int),
    "ls3d_upsample_sum": ([P, P, P, I, I, I, I, I, I, P, P, P], ctypes.c_int),
    "ls3d_normalize_images_u8": ([P, L, P, P, P, I, P], ctypes.c_int),
    "ls3d_upsample_sum_f16": ([P, P, P, I, I, I, I, I, I, P, P, P], ctypes.c_int),
    "ls3d_conv_f16_smem_bytes": ([I, I, I, PL], ctypes.c_int),
    "ls3d_conv_f16_packed_bytes": ([I, I, I, PL], ctypes.c_int),
    "ls3d_conv_f16_pack": ([P, I, I, I, P, P], ctypes.c_int),
    "ls3d_conv_f16": ([P, P, P, P, P, I, I, I, I, I, I, I, P], ctypes.c_int),
    "ls3d_conv_f16_dual": ([P, P, P, P, P, P, I, I, I, I, I, I, I, I, P], ctypes.c_int),
    "ls3d_conv_f16_split_supported": ([I, I, I, I, ctypes.POINTER(ctypes.c_int32)], ctypes.c_int),
    "ls3d_conv_f16_pack_split": ([P, I, I, I, P, P], ctypes.c_int),
    "ls3d_conv_f16_ex": ([ctypes.POINTER(ConvArgs), P], ctypes.c_int),
    "ls3d_conv_f16_multi": ([ctypes.POINTER(ConvArgs), ctypes.POINTER(ConvPass), I, P], ctypes.c_int),
    "ls3d_conv_f16_kb": ([ctypes.POINTER(ConvArgs), ctypes.POINTER(ConvPass), I, P], ctypes.c_int),
    "ls3d_conv_f16_kb_supported": ([I, I, I, I, I, I, L, ctypes.POINTER(ctypes.c_int32)], ctypes.c_int),
    "ls3d_conv_f16_ex_supported": ([I, I, I, I, I, I, ctypes.POINTER(ctypes.c_int32)], ctypes.c_int),
    "ls3d_conv_f16_pack_ex": ([P, I, I, I, I, I, P, P], ctypes.c_int),
    "ls3d_pad3_f16": ([P, L, P, P], ctypes.c_int),
    "ls3d_cast_f32": ([P, P, L, P], ctypes.c_int),
    "ls3d_conv_f16_dual_smem_bytes": ([I, I, I, PL], ctypes.c_int),
    "ls3d_upsample_sum_dual": ([P, P, P, I, I, I, I, I, I, P, P, P, P], ctypes.c_int),
    "ls3d_cast_f16": ([P, P, L, P], ctypes.c_int),
    "ls3d_conv3x3_f16_smem_bytes": ([I, I, PL], ctypes.c_int),
    "ls3d_conv3x3_f16_packed_bytes": ([I, I, PL], ctypes.c_int),
    "ls3d_conv3x3_f16_pack": ([P, I, I, P, P], ctypes.c_int),
    "ls3d_conv3x3_f16": ([P, P, P, P, P, I, I, I, I, I, I, P], ctypes.c_int),
    "ls3d_class_embed_workspace_bytes": ([I, I, I, I, PL], ctypes.c_int),
    "ls3d_class_embed": ([P, I, I, P, I, I, I, P, I, I, P, P, P], ctypes.c_int),
    "ls3d_class_tokens": ([P, I, P, I, I, I, P, I, I, I, P, P, P, P], ctypes.c_int),
}
EXPORTS = ["ls3d_gather_gemm"] + list(_SIGNATURES)


def host_i32(vals):
    return (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])


def host_f32(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


# C-ABI calls since the last COUNTERS.clear(), and the (lower-bound) number of kernels each call launches
COUNTERS = {}
KERNELS_PER_CALL = {"ls3d_gather_gemm": 1, "ls3d_gemm_pack_bf16x3": 1, "ls3d_tile_plan_build": 1, "ls3d_voxelize": 8, "ls3d_vfe_descriptor": 1, "ls3d_vfe_token_attn": 1,
                    "ls3d_vfe_token_max": 1, "ls3d_grid_build": 4, "ls3d_grid_build_strided": 4, "ls3d_grid_enumerate": 1,
                    "ls3d_rulebook_gather": 1, "ls3d_rulebook_scatter": 1, "ls3d_three_nn_grid": 3, "ls3d_three_nn": 1,
                    "ls3d_three_interpolate": 1, "ls3d_frame_offsets": 1, "ls3d_sample_image_features": 1, "ls3d_project_points": 1, "ls3d_project_points_global": 1,
                    "ls3d_resize_images_u8": 1, "ls3d_upsample_sum": 1, "ls3d_upsample_sum_f16": 1,
                    "ls3d_conv3x3_f16": 1, "ls3d_conv3x3_f16_pack": 1, "ls3d_conv_f16": 1, "ls3d_conv_f16_dual": 1, "ls3d_upsample_sum_dual": 1, "ls3d_cast_f16": 1, "ls3d_conv_f16_pack": 1, "ls3d_conv_f16_pack_split": 1, "ls3d_conv_f16_ex": 1, "ls3d_conv_f16_multi": 1, "ls3d_conv_f16_kb": 1, "ls3d_conv_f16_pack_ex": 1, "ls3d_pad3_f16": 1, "ls3d_cast_f32": 1, "ls3d_normalize_images_u8": 1, "ls3d_token_attention": 1, "ls3d_sffm_decoder": 1, "ls3d_class_embed": 4,
                    "ls3d_class_tokens": 1}


def kernel_launches():
    return int(sum(KERNELS_PER_CALL.get(k, 0) * v for k, v in COUNTERS.items()))


def snapshot():
    return dict(COUNTERS)


def add_replay(before, after):
    """A CUDA-graph replay re-launches the kernels captured between two snapshots without passing through ctypes again."""
    for k, v in after.items():
        d = v - before.get(k, 0)
        if d > 0:
            COUNTERS[k] = COUNTERS.get(k, 0) + d


def check(status, what):
    if status != 0:
        raise RuntimeError(f"{what} failed with status {status}")
    COUNTERS[what] = COUNTERS.get(what, 0) + 1


def gather_gemm(args: GemmArgs):
    check(lib().ls3d_gather_gemm(ctypes.byref(args), stream_ptr()), "ls3d_gather_gemm")
