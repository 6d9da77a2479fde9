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


def declared_prototypes():
    """name -> list of parameter declarations of every VPK_API prototype in include/vpk.h"""
    text = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "vpk.h")).read(), flags=re.S)
    protos = {}
    for name, params in re.findall(r"VPK_API\s+[\w\s\*]+?\b(vpk_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = " ".join(params.split())
        protos[name] = [] if params in ("", "void") else [q.strip() for q in params.split(",")]
    return protos


def test_binding_declares_every_parameter_of_every_prototype():
    """ctypes accepts undeclared extra arguments of a cdecl function (no conversion check); every binding must
    therefore declare exactly the parameters of its prototype, pointers as pointers."""
    lib = _lib.load()
    protos = declared_prototypes()
    assert sorted(protos) == sorted(_lib.SYMBOLS)
    for name, params in protos.items():
        at = getattr(lib, name).argtypes
        assert at is not None, "%s: argtypes not declared" % name
        assert len(at) == len(params), "%s: binding declares %d parameters, vpk.h %d" % (name, len(at), len(params))
        for t, decl in zip(at, params):
            is_ptr = "*" in decl or "[" in decl
            ctypes_ptr = t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or issubclass(t, ctypes._Pointer)
            assert is_ptr == ctypes_ptr, "%s: parameter %r bound as %r" % (name, decl, t)
            if "double" in decl and not is_ptr:
                assert t is ctypes.c_double, (name, decl)


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
