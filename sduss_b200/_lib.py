"""ctypes binding of libsduss_b200.so (the C ABI declared in include/sduss_b200.h).

There is no fallback: if the library is missing the import of any product module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SDUSS_B200_LIB: kernel-development hook (tools/build_variant.py) to A/B a differently compiled
# build of the same library; it is never a fallback, the named file must exist.
LIB_PATH = os.environ.get("SDUSS_B200_LIB") or os.path.join(_HERE, "libsduss_b200.so")

OK = 0
ERR_INVALID = 10001
ERR_DRIVER = 10002
ERR_UNSUPPORTED = 10003

c_int = ctypes.c_int
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p
c_i32p = ctypes.POINTER(ctypes.c_int32)


class EpilogueDesc(ctypes.Structure):
    """Mirror of B200EpilogueDesc (include/sduss_b200.h)."""
    _fields_ = [
        ("C", c_void_p), ("ldc", ctypes.c_int32), ("out_fp32", ctypes.c_int32),
        ("bias", c_void_p),
        ("resid", c_void_p), ("ldr", ctypes.c_int32),
        ("gate", c_void_p), ("ldg", ctypes.c_int32),
        ("row_group", c_void_p),
        ("rowvec", c_void_p), ("ldv", ctypes.c_int32),
        ("rms_wq", c_void_p), ("rms_wk", c_void_p),
        ("rms_q_cols", ctypes.c_int32), ("rms_k_cols", ctypes.c_int32),
        ("rms_eps", c_float), ("q_scale", c_float),
        ("act", ctypes.c_int32),
        ("row_mask", c_void_p), ("row_mask_shift", ctypes.c_int32),
        ("stats_out", c_void_p),
        ("ln_stats", c_void_p), ("ln_colsum", c_void_p),
        ("ln_rowpart", c_void_p), ("ln_nparts", ctypes.c_int32), ("ln_eps", c_float),
        ("rowpart_out", c_void_p), ("w_static", ctypes.c_int32), ("row_mask_scale", ctypes.c_int32),
    ]


ACT_DEFAULT, ACT_GELU_TANH, ACT_GELU_ERF, ACT_QUICK_GELU = 0, 1, 2, 3


class AttnExtra(ctypes.Structure):
    """Mirror of B200AttnExtra: causal mask / relative position bias of the text encoders."""
    _fields_ = [("causal", ctypes.c_int32), ("rel_len", ctypes.c_int32), ("rel_bias", c_void_p),
                ("rel_ld", ctypes.c_int32), ("q_mask_shift", ctypes.c_int32), ("q_mask", c_void_p),
                ("bounded_logits", ctypes.c_int32)]


class Forest(ctypes.Structure):
    """Mirror of B200Forest: a flattened RandomForest in device memory (patch cache, row f-3)."""
    _fields_ = [("feature", c_void_p), ("threshold", c_void_p), ("left", c_void_p), ("right", c_void_p),
                ("value", c_void_p), ("roots", c_void_p), ("n_trees", ctypes.c_int32)]


class AttnSource(ctypes.Structure):
    """Mirror of B200AttnSource (include/sduss_b200.h)."""
    _fields_ = [
        ("q", c_void_p), ("ldq", ctypes.c_int32), ("q_col", ctypes.c_int32), ("q_rows", ctypes.c_int32),
        ("k", c_void_p), ("ldk", ctypes.c_int32), ("k_col", ctypes.c_int32),
        ("v", c_void_p), ("ldv", ctypes.c_int32), ("v_col", ctypes.c_int32),
        ("kv_rows", ctypes.c_int32),
        ("out", c_void_p), ("ldo", ctypes.c_int32), ("o_col", ctypes.c_int32),
    ]


class LatentRef(ctypes.Structure):
    """Mirror of B200LatentRef (include/sduss_b200.h): one request of a gather / step launch."""
    _fields_ = [
        ("src", c_void_p), ("dst", c_void_p), ("elems", ctypes.c_int64),
        ("off_a", ctypes.c_int64), ("off_b", ctypes.c_int64),
        ("sigma", c_float), ("sigma_next", c_float),
    ]


DT_BF16, DT_F16, DT_F32 = 0, 1, 2


class B200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m sduss_b200.build` "
            "(nvcc, sm_100a). sduss_b200 has no CPU or PyTorch fallback.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

# name -> argtypes; every function returns int status.
SIGNATURES = {
    "b200_version": [],
    "b200_sm_count": [],
    "b200_attn_rows_per_item": [],
    "b200_gemm_bf16": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                       ctypes.POINTER(EpilogueDesc), c_void_p],
    "b200_attn_build_schedule": [c_void_p, c_int, c_int, c_void_p, ctypes.POINTER(c_int)],
    "b200_attn_varlen_bf16": [ctypes.POINTER(AttnSource), ctypes.POINTER(AttnSource), c_void_p,
                              c_void_p, c_int, c_void_p, c_int, c_float, c_void_p],
    "b200_attn_varlen_ex": [ctypes.POINTER(AttnSource), ctypes.POINTER(AttnSource), c_void_p,
                            c_void_p, c_int, c_void_p, c_int, c_float, ctypes.POINTER(AttnExtra),
                            c_void_p],
    "b200_embed_rows_bf16": [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                             c_void_p],
    "b200_rmsnorm_bf16": [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p],
    "b200_layernorm_mod_bf16": [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                                c_int, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "b200_attn_cross_short_bf16": [ctypes.POINTER(AttnSource), ctypes.POINTER(AttnSource), c_void_p, c_int, c_int,
                                   c_int, c_int, c_float, c_void_p, c_int, c_void_p],
    "b200_patch_mask_bf16": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(Forest), c_int, c_int,
                             c_void_p, c_void_p],
    "b200_patch_mask_ex": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(Forest), c_int, c_int,
                           c_void_p, c_int, c_void_p, c_void_p],
    "b200_row_stats_bf16": [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p],
    "b200_silu_bf16": [c_void_p, c_void_p, ctypes.c_longlong, c_void_p],
    "b200_timestep_embedding": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p],
    "b200_sd3_patchify": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                          c_void_p],
    "b200_sd3_unpatchify": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                            c_void_p],
    "b200_gather_latents": [ctypes.POINTER(LatentRef), c_int, c_int, c_int, c_void_p, c_void_p],
    "b200_cfg_scheduler_step": [c_void_p, c_int, ctypes.POINTER(LatentRef), c_int, c_int, c_float,
                                c_int, c_int, c_void_p],
    "b200_write_f32": [c_void_p, ctypes.POINTER(c_float), c_int, c_void_p],
    "b200_gather_rows": [c_void_p, ctypes.c_longlong, ctypes.POINTER(c_void_p), c_int,
                         ctypes.c_longlong, c_void_p],
    "b200_conv3x3_encode_maps": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "b200_conv3x3_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                          c_int, c_void_p, c_int, c_int, ctypes.POINTER(EpilogueDesc), c_void_p],
    "b200_groupnorm_nhwc_bf16": [c_void_p, c_int, ctypes.c_longlong, c_int, c_int, c_float,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                 c_int, c_void_p, c_void_p],
    "b200_groupnorm_from_conv_stats": [c_void_p, c_int, ctypes.c_longlong, c_int, c_int, c_float,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_void_p, c_int, c_void_p, c_void_p],
    "b200_pack_im2col3x3": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "b200_scatter_nchw": [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "b200_upsample2x_nhwc": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                             c_int, c_void_p],
    "b200_copy_cols_bf16": [c_void_p, c_int, c_void_p, c_int, ctypes.c_longlong, c_int, c_void_p],
    "b200_split_patches": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "b200_concat_patches": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                            c_void_p],
    "b200_latent_affine": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                           c_void_p, c_void_p],
    "b200_softmax_rows": [c_void_p, ctypes.c_longlong, c_int, c_int, c_float, c_void_p,
                          ctypes.c_longlong, c_void_p],
}


def _bind():
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.b200_groupnorm_workspace_bytes.argtypes = [ctypes.c_longlong, c_int]
    lib.b200_groupnorm_workspace_bytes.restype = ctypes.c_longlong
    lib.b200_patch_mask_workspace_bytes.argtypes = [c_int]
    lib.b200_patch_mask_workspace_bytes.restype = ctypes.c_longlong
    lib.b200_attn_cross_short_max_keys.argtypes = []
    lib.b200_attn_cross_short_max_keys.restype = c_int
    lib.b200_attn_workspace_bytes.argtypes = []
    lib.b200_attn_workspace_bytes.restype = ctypes.c_longlong
    lib.b200_conv3x3_maps_bytes.argtypes = [c_int]
    lib.b200_conv3x3_maps_bytes.restype = ctypes.c_longlong


_bind()


def check(status, what):
    if status != OK:
        raise B200Error(f"{what} failed with status {status}")
