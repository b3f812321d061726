"""CPU-only checks of the C ABI: the library loads, exports every symbol include/frankb200.h declares, and refuses
to create a context without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'frankb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(fb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from frank_b200 import _lib
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/frankb200.h but not exported"
    # and the Python binding declares a signature for each of them
    assert set(declared) == set(_lib.exported_symbols())
    assert lib.fb_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from frank_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.Context(0)
    from frank_b200.geometry import FixedGeometry
    from frank_b200.hankel import DiscreteHankelTransform
    from frank_b200.statistical_models import VisibilityMapping
    import numpy as np
    vm = VisibilityMapping(DiscreteHankelTransform(1e-5, 20), FixedGeometry(0, 0), verbose=False)
    with pytest.raises(RuntimeError):
        vm.map_visibilities(np.ones(4), np.ones(4), np.ones(4, dtype=complex), np.ones(4))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under frank_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'frank_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, fn)).read()
                assert 'oracle' not in txt.replace('frank_oracle_unused', ''), os.path.join(dirpath, fn)


def test_binding_arity_matches_header():
    """Every ctypes signature in frank_b200/_lib.py has as many arguments as the prototype in include/frankb200.h,
    with pointers, 64-bit counts, ints and doubles in the same positions (guards against ABI drift)."""
    from frank_b200 import _lib
    src = open(os.path.join(ROOT, 'include', 'frankb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = dict(re.findall(r'\b(fb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', src))
    assert set(protos) == set(_lib.exported_symbols())

    def kind(arg):
        arg = arg.strip()
        if arg in ('void', ''):
            return None
        if '*' in arg:
            return 'p'
        if 'int64_t' in arg:
            return 'l'
        if arg.startswith('double'):
            return 'd'
        return 'i'
    ckind = {ctypes.c_void_p: 'p', ctypes.c_int64: 'l', ctypes.c_longlong: 'l', ctypes.c_int: 'i', ctypes.c_double: 'd',
             ctypes.c_char_p: 'p'}
    for name, args in protos.items():
        want = [k for k in (kind(a) for a in args.split(',')) if k]
        argtypes, _ = _lib._SIGNATURES[name]
        got = [ckind.get(t, 'p') for t in argtypes]          # POINTER(...) types count as pointers
        assert got == want, f"{name}: header {want} vs binding {got}"
