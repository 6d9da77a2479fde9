"""CPU-side checks of the drop-in boundary: libvpk.so loads without a GPU and
exports every symbol include/vpk.h declares."""
import ctypes
import os
import re

from vanishing_points_2017_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vpk.h")).read()
    return sorted(set(re.findall(r"VPK_API\s+[\w\s\*]+?\b(vpk_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.vpk_abi_version() == 1


def test_em_config_defaults_match_reference_kwargs():
    cfg = _lib.em_config()
    # vp_localisation.py:168-172
    assert (cfg.num_iter, cfg.num_init_vp, cfg.split_merge_freq, cfg.num_min_lines) == (100, 25, 10, 3)
    assert (cfg.do_merge, cfg.do_split, cfg.do_iterations, cfg.use_weights) == (1, 1, 1, 1)
    assert cfg.wbias == 1 and cfg.merge_thresh == 1e-3 and cfg.outlier_thresh == 1.96 ** 2
    assert cfg.final_convergence == 5e-3 and cfg.s_thresh == 1e-200


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(_lib.VpkError):
        _lib.Context(0)
