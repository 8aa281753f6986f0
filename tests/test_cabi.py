"""CPU: the C-ABI library loads and exports every symbol include/aemcarl_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aemcarl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aem_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from aemcarl_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert lib.aem_version() == 100


def test_error_reporting_without_gpu():
    from aemcarl_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.aem_create(0, None) == 1                      # AEM_ERR_INVALID: out is null
    assert b"null" in lib.aem_last_error()
    import torch
    if not torch.cuda.is_available():
        rc = lib.aem_create(0, ctypes.byref(h))
        assert rc != 0 and len(lib.aem_last_error()) > 0     # fails loudly, no silent CPU path
    assert lib.aem_launch_count(None) == 0


def test_struct_layout_matches_header():
    from aemcarl_b200._lib import EnvParams
    assert ctypes.sizeof(EnvParams) == 10 * 8 + 8 + 8
    from oracle.pyoracle import EnvParams as OP
    assert ctypes.sizeof(OP) == ctypes.sizeof(EnvParams)


def test_ctypes_signatures_have_the_argument_count_of_the_header():
    """every binding in aemcarl_b200/_lib.py passes as many arguments as the prototype in include/aemcarl_b200.h declares"""
    from aemcarl_b200 import _lib
    src = open(os.path.join(ROOT, "include", "aemcarl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = dict(re.findall(r"\b(aem_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(protos) == set(_lib.SYMBOLS)
    for name, args in protos.items():
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(_lib.SYMBOLS[name][1]), (name, n, len(_lib.SYMBOLS[name][1]))
