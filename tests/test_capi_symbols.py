"""The C-ABI library loads and exports every symbol include/at3d_b200.h declares (no compute calls:
this test runs without a GPU), and the compute entry points fail loudly without a CUDA device."""
import ctypes as C
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from at3d_b200 import _lib
    return _lib.lib()


def declared_symbols():
    h = open(os.path.join(ROOT, 'include', 'at3d_b200.h')).read()
    h = re.sub(r'/\*.*?\*/', '', h, flags=re.S)
    return sorted(set(re.findall(r'\b(at3d_[a-z0-9_]+)\s*\(', h)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    dll = C.CDLL(os.path.join(ROOT, 'at3d_b200', 'lib', 'libat3d_b200.so'))
    for n in names:
        assert hasattr(dll, n), 'libat3d_b200.so does not export %s' % n


def test_binding_table_matches_header():
    from at3d_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_version_and_no_cpu_fallback(lib):
    assert b'sm_100a' in lib.at3d_b200_version()
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present: the no-device error path is not reachable')
    from at3d_b200 import backend as B
    from at3d_b200._lib import At3dError
    with pytest.raises(At3dError) as e:
        B.ylmall(False, 0.5, 0.1, 7, 7, 1, 64)
    assert e.value.code == 4 and 'no CPU fallback' in e.value.msg


def test_oracle_is_not_reachable_from_the_product():
    """Nothing under at3d_b200/ may import or load the oracle (test infrastructure)."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, 'at3d_b200')):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f), errors='replace').read()
                if re.search(r'^\s*(import|from)\s+oracle_lib|libshdom_oracle\.so|#include\s*[<"].*oracle', txt, re.M):
                    bad.append(f)
    assert bad == [], bad
