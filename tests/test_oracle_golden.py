"""Oracle vs committed golden vectors (tests/golden/*.npz, generated from the real reference by
tests/golden/make_golden.py).  Needs neither /root/reference nor oracle/_ref: runs anywhere."""
import ctypes as C
import os

import numpy as np
import pytest

from common import LUMA_PU, LUMA_CU, P, I, IP, load_oracle, ptr, pixel_dtype
from me_helpers import run_both

HERE = os.path.dirname(os.path.abspath(__file__))


class _OracleAsRef:
    """run_both() wants a 'reference' object; here the oracle plays both roles and the golden file holds the truth."""


@pytest.mark.parametrize("depth", [8, 10])
def test_oracle_matches_golden(depth):
    g = np.load(os.path.join(HERE, "golden", "primitives_%d.npz" % depth))
    O = load_oracle(depth)
    dt = pixel_dtype(depth)
    a, b = g["cmp_a"], g["cmp_b"]
    for pu in (1, 2, 4, 6, 8, 14, 17, 22):
        w, h = LUMA_PU[pu]
        assert O.orc_sad(ptr(a), IP(64), ptr(b, 5), IP(96), w, h) == int(g["sad_%d" % pu])
        assert O.orc_satd(ptr(a), IP(64), ptr(b, 5), IP(96), w, h) == int(g["satd_%d" % pu])
    for cu, n in enumerate(LUMA_CU):
        assert O.orc_sa8d(ptr(a), IP(64), ptr(b, 5), IP(96), n, n) == int(g["sa8d_%d" % cu])
        assert O.orc_sse_pp(ptr(a), IP(64), ptr(b, 5), IP(96), n, n) == int(g["sse_%d" % cu])
        assert O.orc_psy_cost_pp(ptr(a), IP(64), ptr(b, 5), IP(96), n) == int(g["psy_%d" % cu])
        assert O.orc_var(ptr(a), IP(64), n) == int(g["var_%d" % cu])
    src, ssrc = g["ip_src"], g["ip_ssrc"]
    s0 = 8 * 64 + 8
    for ci in range(1, 4):
        for nm, fn, short_in, short_out in (("luma_hpp", O.orc_interp_hpp, 0, 0), ("luma_vpp", O.orc_interp_vpp, 0, 0), ("luma_vps", O.orc_interp_vps, 0, 1),
                                            ("luma_vsp", O.orc_interp_vsp, 1, 0), ("luma_vss", O.orc_interp_vss, 1, 1)):
            d = np.zeros((16, 16), np.int16 if short_out else dt)
            fn(ptr(ssrc if short_in else src, s0), IP(64), ptr(d), IP(16), ci, 8, 16, 16)
            assert np.array_equal(d, g["%s_%d" % (nm, ci)]), nm
        d = np.zeros((23, 16), np.int16)
        O.orc_interp_hps(ptr(src, s0), IP(64), ptr(d), IP(16), ci, 1, 8, 16, 16)
        assert np.array_equal(d, g["luma_hps_ext_%d" % ci])
        for cj in range(1, 4):
            d = np.zeros((16, 16), dt)
            O.orc_interp_hvpp(ptr(src, s0), IP(64), ptr(d), IP(16), ci, cj, 8, 16, 16)
            assert np.array_equal(d, g["luma_hvpp_%d%d" % (ci, cj)])
    for cu, n in enumerate(LUMA_CU[:4]):
        res = g["tr_res_%d" % n]
        co = np.zeros(n * n, np.int16); O.orc_dct(ptr(res), ptr(co), IP(n), n)
        assert np.array_equal(co, g["dct_%d" % n])
        back = np.zeros((n, n), np.int16); O.orc_idct(ptr(co), ptr(back), IP(n), n)
        assert np.array_equal(back, g["idct_%d" % n])
        qbits, add = [int(x) for x in g["quant_par_%d" % n]]
        qc = np.full(n * n, 18396, np.int32); du = np.zeros(n * n, np.int32); q = np.zeros(n * n, np.int16)
        O.orc_quant.restype = C.c_uint32
        ns = O.orc_quant(ptr(co), ptr(qc), ptr(du), ptr(q), qbits, add, n * n)
        assert ns == int(g["quant_ns_%d" % n]) and np.array_equal(q, g["quant_%d" % n]) and np.array_equal(du, g["quant_du_%d" % n])
        dq = np.zeros(n * n, np.int16); O.orc_dequant_normal(ptr(q), ptr(dq), n * n, 57 << 5, 6 - (15 - depth - (cu + 2)))
        assert np.array_equal(dq, g["dequant_%d" % n])
    co = np.zeros(16, np.int16); O.orc_dst4(ptr(g["tr_res_4"]), ptr(co), IP(4)); assert np.array_equal(co, g["dst4"])
    for cu, n in enumerate(LUMA_CU[:4]):
        nb = g["intra_nb_%d" % n]
        f = np.zeros(4 * n + 1, dt); O.orc_intra_filter(ptr(nb), ptr(f), n); assert np.array_equal(f, g["intra_filt_%d" % n])
        for m in range(35):
            out = np.zeros((n, n), dt)
            O.orc_intra_pred(ptr(out), IP(n), ptr(nb), m, 1 if n <= 16 else 0, n)
            assert np.array_equal(out, g["intra_pred_%d" % n][m]), (n, m)


class _NoRef:
    def __init__(self, lam):
        self._lam = lam

    def x265ref_lambda(self, qp):
        return self._lam

    class _Fn:
        argtypes = None

        def __call__(self, *a):
            return 0
    x265ref_motion_estimate = _Fn()
    x265ref_motion_estimate_chroma = _Fn()


@pytest.mark.parametrize("depth", [8, 10])
def test_oracle_me_matches_golden(depth):
    """Regenerates the seeded ME jobs of make_golden.py and compares the oracle's result with the stored
    result of the real MotionEstimate."""
    from frame_helpers import lambda_for
    g = np.load(os.path.join(HERE, "golden", "primitives_%d.npz" % depth))
    O = load_oracle(depth)
    R = _NoRef(lambda_for(30, depth))
    mrng = np.random.default_rng(77 + depth)
    k = 0
    for method in (0, 1, 3):
        for (w, h) in ((8, 8), (16, 16), (32, 16), (64, 64), (8, 4)):
            for subme in (1, 3, 5):
                _, o = run_both(O, R, depth, mrng, w, h, method, subme, 0, True, 57 if method == 3 else 16)
                row = g["me_results"][k]
                assert list(row[:4]) == [method, w, h, subme]
                assert list(o) == [int(x) for x in row[4:7]], (method, w, h, subme, o, row)
                k += 1


@pytest.mark.parametrize("depth", [8, 10])
def test_oracle_me_chroma_matches_golden(depth):
    """The chroma-SATD term of subpelCompare (motion.cpp:1601-1661): the oracle against the stored results of the real
    MotionEstimate (tests/golden/me_chroma_*.npz, make_golden.py::gen_chroma), inputs regenerated from the seed."""
    from frame_helpers import lambda_for
    from me_helpers import run_both_chroma, CHROMA_CASES
    g = np.load(os.path.join(HERE, "golden", "me_chroma_%d.npz" % depth))["me_results"]
    O = load_oracle(depth)
    R = _NoRef(lambda_for(30, depth))
    rng = np.random.default_rng(177 + depth)
    assert len(g) == len(CHROMA_CASES)
    for k, (method, w, h, subme) in enumerate(CHROMA_CASES):
        _, o = run_both_chroma(O, R, depth, rng, w, h, method, subme, True, 57 if method == 3 else 16)
        assert list(g[k][:4]) == [method, w, h, subme]
        assert list(o) == [int(x) for x in g[k][4:7]], (method, w, h, subme, o, g[k])
