"""CPU: the C-ABI library loads, exports every symbol include/rd_b200.h declares, and the ctypes signatures
in rd_b200/lib.py agree with the header parameter lists (no compute calls without a GPU)."""
import re

import pytest

import rd_b200.lib as L


def _header_decls():
    with open(L.HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct rd_conv_desc \{.*?\} rd_conv_desc;", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(rd_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, params = m.group(1), m.group(2)
        plist = [p.strip() for p in params.split(",") if p.strip() and p.strip() != "void"]
        decls[name] = plist
    return decls


def _ctype_of(param: str):
    if "*" in param or "rd_stream" in param:
        return L.P
    if "int64_t" in param:
        return L.L
    if "float" in param:
        return L.F
    return L.I


def test_library_exports_every_header_symbol():
    lib = L.load()
    decls = _header_decls()
    assert len(decls) >= 50
    for name in decls:
        assert hasattr(lib, name), "librd_b200.so does not export " + name
    assert lib.rd_abi_version() == 1
    assert sorted(decls) == L.header_symbols()


def test_ctypes_signatures_match_header():
    decls = _header_decls()
    for name, sig in L._SIGS.items():
        assert name in decls, name
        params = decls[name]
        assert _ctype_of(params[0]) is L.P and "rd_ctx" in params[0], name
        want = [_ctype_of(p) for p in params[1:]]
        assert len(want) == len(sig), "%s: header has %d params after ctx, binding has %d" % (name, len(want), len(sig))
        for k, (a, b) in enumerate(zip(want, sig)):
            assert a is b, "%s: param %d (%s) header %s vs binding %s" % (name, k + 1, params[k + 1], a, b)
    bound = set(L._SIGS) | {"rd_abi_version", "rd_ctx_create", "rd_ctx_destroy", "rd_last_error", "rd_launch_count",
                            "rd_last_conv_algo", "rd_norm_partial_chunks", "rd_spade_bwd_workspace", "rd_mix_job_blocks", "rd_mixf_job_blocks", "rd_metrics_recon_tiles", "rd_wgrad_tma_plan"}
    assert bound == set(decls), (set(decls) - bound, bound - set(decls))


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rd_b200.kernels as K
    x = torch.zeros(1, 4, 4, 8)
    with pytest.raises(L.RdError):
        K.lrelu_fwd(x, torch.empty_like(x), 0.2)
    with pytest.raises(L.RdError):
        L.get_ctx(0)


def test_wgrad_tma_split_plan_properties():
    """rd_wgrad_tma_plan (host helper, no GPU): the per-launch split of the TMA weight-gradient kernel over the step's layer shapes
    at several batch sizes and SM counts.  Every X box and every pixel tile is covered exactly once with no empty CTA, a group is cut into at most
    ceil(tiles / 8) split-K chunks, the accumulator fits the tensor memory, and the launch never has fewer CTAs than the plain
    "two waves" rule unless it still fills about two waves."""
    import rd_b200.kernels as K
    shapes = [  # H, W, cin, cout, k, stride, images per B, groups per B-independent
        (40, 48, 128, 256, 3, 1, 16, 16), (40, 48, 128, 64, 3, 1, 16, 16), (20, 24, 128, 256, 3, 1, 16, 16), (20, 24, 128, 128, 3, 1, 16, 16),
        (10, 12, 128, 256, 3, 1, 16, 16), (5, 6, 128, 128, 3, 1, 16, 16), (5, 6, 16, 128, 3, 1, 16, 16), (10, 12, 16, 128, 3, 1, 16, 16),
        (40, 48, 256, 64, 3, 1, 4, 4), (20, 24, 512, 128, 3, 1, 4, 4), (10, 12, 256, 256, 3, 1, 4, 4), (10, 12, 256, 256, 4, 2, 4, 4),
        (20, 24, 128, 256, 4, 2, 4, 4), (40, 48, 64, 128, 4, 2, 4, 4), (80, 96, 32, 64, 4, 2, 4, 4), (160, 192, 16, 32, 4, 2, 4, 4),
        (160, 192, 16, 16, 3, 2, 4, 4), (80, 96, 16, 32, 3, 2, 4, 4), (40, 48, 32, 64, 3, 2, 4, 4), (20, 24, 64, 128, 3, 2, 4, 4),
        (10, 12, 128, 128, 3, 2, 4, 4)]
    checked = changed = 0
    for B in (1, 2, 3, 4, 8, 16):
        for sm in (132, 148):
            for (H, W, cin, cout, k, s, ipb, g) in shapes:
                d = K.conv_desc(ipb * B, H, W, cin, cout, k, k, s, 1, g, L.RD_BF16)
                p = K.wgrad_tma_plan(d, sm)
                if p is None:
                    continue
                checked += 1
                tag = (B, sm, H, W, cin, cout, k, s, p)
                assert p["xsplits"] * p["xb_per_cta"] >= p["xb_total"] > (p["xsplits"] - 1) * p["xb_per_cta"], tag
                assert p["chunks_per_group"] * p["chunk_tiles"] >= p["tiles_per_group"] > (p["chunks_per_group"] - 1) * p["chunk_tiles"], tag
                assert p["chunks_per_group"] <= -(-p["tiles_per_group"] // 8), tag          # split-K chunks of about 8 pixel tiles or more
                bi = 64 if cin % 64 == 0 else 32 if cin % 32 == 0 else 16
                cols = p["xb_per_cta"] * bi if cout >= 128 else -(-p["xb_per_cta"] * bi // 128) * cout
                assert cols + 16 <= 512, tag
                gy = cout // 128 if cout >= 128 else 1
                assert p["ctas"] == p["chunks_per_group"] * p["xsplits"] * gy * g, tag
                assert p["ctas"] >= min(p["ctas_two_waves"], 2 * sm - sm // 8), tag
                changed += p["ctas"] != p["ctas_two_waves"]
                # the same query twice gives the same split (cached per shape)
                assert K.wgrad_tma_plan(d, sm) == p
    assert checked >= 150 and changed >= 20, (checked, changed)
    # the launch the round-2 capture was taken on: 320 CTAs (2.16 waves, last X split half empty) before
    p = K.wgrad_tma_plan(K.conv_desc(256, 40, 48, 128, 256, 3, 3, 1, 1, 16, L.RD_BF16), 148)
    assert p["ctas_two_waves"] == 320 and p["xb_per_cta"] == 3 and p["xsplits"] == 6 and p["ctas"] % 144 == 0, p

