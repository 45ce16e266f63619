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
                            "rd_last_conv_algo", "rd_norm_partial_chunks", "rd_spade_bwd_workspace", "rd_mix_job_blocks", "rd_mixf_job_blocks", "rd_metrics_recon_tiles"}
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
