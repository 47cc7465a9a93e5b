"""GPU parity: the CUDA per-call primitive table (x265cu_get_primitive, the entries
setupCudaPrimitives() installs) against the oracle, field by field -- the same checks that pin the
oracle to the reference C table in test_oracle_vs_ref.py (reference TestBench strategy,
source/test/testbench.cpp:155-233).  Bit-exact."""
import numpy as np
import pytest

import table_checks
from common import load_oracle

pytestmark = pytest.mark.gpu

DEPTHS = [8, 10]
KINDS = ["rand", "min", "max"]


@pytest.fixture(scope="module")
def cu():
    import x265_b200
    return x265_b200.load()


def cuda_getter(cu, depth):
    def get(name, restype, argtypes, i=0, j=0, k=0):
        return cu.primitive(depth, name, restype, argtypes, i, j, k)
    return get


@pytest.mark.parametrize("depth", DEPTHS)
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("cls", ["pixelcmp", "blockops", "interp", "transforms", "intra"])
def test_table(cu, depth, kind, cls):
    getattr(table_checks, "check_" + cls)(cuda_getter(cu, depth), load_oracle(depth), depth, kind)


@pytest.mark.parametrize("depth", DEPTHS)
def test_ads(cu, depth):
    table_checks.check_ads(cuda_getter(cu, depth), load_oracle(depth), depth)


def test_native_library_loaded(cu):
    # the product path is the CUDA library, in-tree, and it launched kernels
    assert cu.path.endswith("x265_b200/libx265cu.so")
    with open("/proc/self/maps") as f:
        assert "libx265cu.so" in f.read()


@pytest.mark.parametrize("depth", DEPTHS)
def test_integral(cu, depth):
    table_checks.check_integral(cuda_getter(cu, depth), load_oracle(depth), depth)
