"""CPU-only checks: the C-ABI library loads without a GPU and exports every symbol
include/sduss_b200.h declares; ctypes signatures cover the header; argument validation that
does not need a device."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sduss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(b200_\w+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from sduss_b200 import _lib
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/sduss_b200.h but not exported"
    assert _lib.lib.b200_version() == 2


def test_ctypes_signatures_cover_header():
    from sduss_b200 import _lib
    declared = set(_header_functions())
    bound = set(_lib.SIGNATURES) | {"b200_groupnorm_workspace_bytes", "b200_patch_mask_workspace_bytes",
                                    "b200_attn_workspace_bytes", "b200_conv3x3_maps_bytes",
                                    "b200_attn_cross_short_max_keys"}
    assert declared == bound, declared ^ bound


def test_struct_layouts_match_header():
    from sduss_b200 import _lib
    # B200EpilogueDesc: 8B ptr, 2x i32, 2 ptr, i32(+pad), ptr, i32(+pad), ptr, ptr, i32(+pad), 2 ptr, 2 i32, 2 f32
    assert ctypes.sizeof(_lib.EpilogueDesc) == 192 and _lib.EpilogueDesc.w_static.offset == 184 and _lib.EpilogueDesc.act.offset == 112
    assert _lib.EpilogueDesc.ln_rowpart.offset == 160 and _lib.EpilogueDesc.rowpart_out.offset == 176
    assert _lib.EpilogueDesc.ln_stats.offset == 144
    assert _lib.EpilogueDesc.stats_out.offset == 136
    assert _lib.EpilogueDesc.row_mask.offset == 120 and _lib.EpilogueDesc.row_mask_scale.offset == 188
    # B200AttnExtra: 2 i32, ptr, 2 i32, ptr, i32 (+ pad)
    assert ctypes.sizeof(_lib.AttnExtra) == 40 and _lib.AttnExtra.q_mask.offset == 24
    assert _lib.AttnExtra.bounded_logits.offset == 32
    assert _lib.EpilogueDesc.row_group.offset == 56 and _lib.EpilogueDesc.rms_eps.offset == 104
    assert ctypes.sizeof(_lib.AttnSource) == 80
    assert _lib.AttnSource.k.offset == 24 and _lib.AttnSource.out.offset == 64


def test_invalid_arguments_are_rejected_without_a_device():
    from sduss_b200 import _lib
    lib = _lib.lib
    assert lib.b200_gemm_bf16(None, 0, None, 0, 0, 0, 0, 0, None, None) == _lib.ERR_INVALID
    assert lib.b200_silu_bf16(None, None, 8, None) == _lib.ERR_INVALID
    assert lib.b200_groupnorm_workspace_bytes(128, 2) == 2 * 32 * 2 * 4 + 2 * 32 * 2 * 4 + 1024 * 4  # + grid-barrier words
    assert lib.b200_attn_workspace_bytes() == 8 and lib.b200_conv3x3_maps_bytes(3) == 384
    assert lib.b200_patch_mask_workspace_bytes(10) == 10 * (16 * 4 + 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sduss_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f
