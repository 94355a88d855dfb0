"""The C-ABI library loads on a CPU-only box and exports every symbol include/facevae_b200.h declares
(no compute call is made here: there is no GPU in the build container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "facevae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fv_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for needed in ("fv_conv2d", "fv_conv2d_wgrad", "fv_bn_act_fwd", "fv_bn_act_bwd_apply", "fv_reparam_kl_fwd",
                   "fv_recon_loss", "fv_last_error", "fv_version"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    from face_vae_b200 import _lib
    lib = _lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert b"sm_100a" in lib.fv_version()
    # every int-returning declaration has a ctypes signature in the binding
    missing = [s for s in declared_symbols() if s not in _lib.SIGNATURES and s not in _lib._STR and s not in _lib._LL and s not in _lib._PLAIN_INT]
    assert not missing, missing


def test_no_cpu_fallback_in_product_package():
    """The product must not import the oracle nor carry a CPU code path (prompt (3))."""
    pkg = os.path.join(ROOT, "face_vae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src, f"{f} reads the reference tree"


def test_argument_errors_do_not_need_a_gpu():
    from face_vae_b200 import _lib
    lib = _lib.load()
    rc = lib.fv_conv2d(None, None, None, None, None, 0, 1, 8, 8, 16, 16, 16, 3, 3, 1, None)
    assert rc != 0 and b"null pointer" in lib.fv_last_error()
    rc = lib.fv_reparam_kl_fwd(None, None, 0, None, None, None, 1, 7, None)
    assert rc != 0
