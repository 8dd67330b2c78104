"""The C-ABI library loads and exports every symbol include/nbco.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbco.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbco_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import coulomb_oscillators_b200._lib as L
    lib = ctypes.CDLL(L.lib_path)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(L.SYMBOLS) == syms, set(L.SYMBOLS) ^ set(syms)


def test_abi_version_and_defaults():
    import coulomb_oscillators_b200 as nb
    assert nb.lib.nbco_abi_version() == 3
    cfg = nb._lib.default_config()
    # defaults of reference constants.cuh:36-52
    assert (cfg.order, cfg.tree_steps, cfg.coll, cfg.unsort, cfg.max_level) == (3, 8, 1, 1, 0)
    assert cfg.radius == 1.0 and abs(cfg.eps2 - 1e-18) < 1e-24 and cfg.dens_inhom == 1.0
    assert cfg.eps2_d == 1e-18


def test_no_cpu_fallback():
    """without a GPU the product must fail loudly, never compute on the host"""
    import torch
    import coulomb_oscillators_b200 as nb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nb.NbcoError, match="no CUDA device"):
        nb.Context()


def test_product_does_not_link_oracle():
    import coulomb_oscillators_b200._lib as L
    blob = open(L.lib_path, "rb").read()
    assert b"orc_fmm3_kd" not in blob and b"orc2_fmm" not in blob and b"libnbco_oracle" not in blob and b"libnbco_ref" not in blob
