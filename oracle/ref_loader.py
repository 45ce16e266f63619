"""TEST INFRASTRUCTURE (oracle/): loader for the UNMODIFIED reference `src/model.py`.

Only usable where `/root/reference` exists (the build container).  It is used by
`oracle/make_golden.py` to (a) pin `oracle/rd_oracle.py` against the real reference and
(b) write the committed fixtures under `tests/golden/`.  Nothing in the product package,
`bench.py` or the `-m gpu` tests imports this file.

The reference cannot be imported as shipped because `src/util.py:14-29` imports
skimage / matplotlib / nibabel / h5py / nonechucks (absent here); we pre-register empty
stub modules for those names, never touching the reference tree.
"""
import contextlib
import io
import os
import sys
import types

REF_SRC = os.environ.get("RD_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference_model_module():
    """Return the reference `model` module (imported once, stdout silenced)."""
    if "rd_reference_model" in sys.modules:
        return sys.modules["rd_reference_model"]
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REF_SRC)
    try:
        import scipy.misc  # noqa: F401  (util.py:14 `import scipy.misc as sci`)
    except Exception:
        import scipy
        scipy.misc = _stub("scipy.misc")
    sk = _stub("skimage")
    for sub in ("io", "transform", "color", "metrics"):
        setattr(sk, sub, _stub("skimage." + sub))
    sk.measure = _stub("skimage.measure", compare_nrmse=None, compare_psnr=None, compare_ssim=None)
    _stub("matplotlib")
    _stub("nibabel")
    _stub("h5py")
    _stub("nonechucks")
    sys.path.insert(0, REF_SRC)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import model as ref_model  # noqa
    finally:
        sys.path.remove(REF_SRC)
    sys.modules["rd_reference_model"] = ref_model
    return ref_model


def build_reference_model(cfg, device="cpu"):
    """Construct the reference MultimodalModel exactly like src/main_missing.py:71-95."""
    import torch
    ref = load_reference_model_module()
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.MultimodalModel(
            input_size=(cfg["input_height"], cfg["input_width"]),
            modality_num=len(cfg["contrast_list"]), in_num_ch=2 * cfg["block_size"] + 1,
            out_num_ch=cfg["out_num_ch"], s_num_ch=cfg["s_num_ch"], z_size=cfg["z_size"],
            is_cond=cfg["is_cond"], is_discrim_s=cfg["lambda_adv_s"] > 0, is_distri_z=cfg["is_distri_z"],
            s_compact_method=cfg["s_compact_method"], s_sim_method=cfg["s_sim_method"],
            z_sim_method=cfg["z_sim_method"], shared_ana_enc=cfg["shared_ana_enc"],
            shared_mod_enc=cfg["shared_mod_enc"], shared_inp_dec=cfg["shared_inp_dec"],
            device=torch.device(device), input_output_act=cfg["input_output_act"],
            target_output_act=cfg["target_output_act"], target_model_name=cfg["target_model_name"],
            fuse_method=cfg["fuse_method"], others=cfg["others"])
    return m
