"""Pins the lookahead restatement (oracle/oracle_lookahead.c: lowresIntraEstimate + estimateFrameCost /
estimateCUCost) against the REAL classes compiled from /root/reference (Lowres::init,
LookaheadTLD::lowresIntraEstimate, CostEstimateGroup::singleCost) through oracle/_ref.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from common import load_ref, load_oracle, ptr, P, I, IP, pixel_dtype
from frame_helpers import gen_luma, MARGIN_X, MARGIN_Y
from lookahead_helpers import OracleLookahead

TRIPLES = [(0, 1, 1), (0, 2, 1), (0, 2, 2), (0, 3, 1), (0, 3, 2), (0, 3, 3), (1, 3, 2), (2, 3, 3)]


class RefLookahead:
    def __init__(self, R, frames, bframes=3):
        self.R = R
        H, W = frames[0].shape
        self.keep = [np.ascontiguousarray(f) for f in frames]
        arr = (P * len(frames))(*[C.c_void_p(f.ctypes.data) for f in self.keep])
        R.x265ref_la_create.restype = P; R.x265ref_la_create.argtypes = [I, I, I, P, IP, I]
        R.x265ref_la_cost.restype = C.c_int64; R.x265ref_la_cost.argtypes = [P, I, I, I]
        R.x265ref_la_get.argtypes = [P, I, I, I, I, P]; R.x265ref_la_geometry.argtypes = [P, P]
        self.h = R.x265ref_la_create(W, H, len(frames), arr, W, bframes)
        g = np.zeros(7, np.int32); R.x265ref_la_geometry(self.h, ptr(g))
        self.w8, self.h8, self.stride, self.lw, self.lh = [int(x) for x in g[:5]]
        self.ncu = self.w8 * self.h8

    def cost(self, p0, p1, b):
        return int(self.R.x265ref_la_cost(self.h, p0, p1, b))

    def get(self, frame, what, d0=0, d1=0, dtype=np.int32, n=None):
        out = np.zeros(n if n is not None else self.ncu, dtype)
        assert self.R.x265ref_la_get(self.h, frame, what, d0, d1, ptr(out)) == 0
        return out


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("size,noise", [((416, 240), False), ((200, 136), True)])
def test_lookahead(depth, size, noise):
    R = load_ref(depth)
    if R is None:
        pytest.skip("oracle/_ref not built")
    O = load_oracle(depth)
    W, H = size
    frames = [gen_luma(W, H, i, s1=17.0, s2=11.0, bits=depth, noise=noise) for i in range(4)]
    ref = RefLookahead(R, frames)
    orc = OracleLookahead(O, frames, depth)
    assert (ref.w8, ref.h8, ref.stride) == (orc.w8, orc.h8, orc.stride)
    # lowres planes incl. margins (Lowres::init)
    rows = ref.lh + 2 * MARGIN_Y
    for k in range(4):
        pl = ref.get(1, 7, k, 0, pixel_dtype(depth), rows * ref.stride).reshape(rows, ref.stride)
        vw = ref.lw + 2 * MARGIN_X
        assert np.array_equal(pl[:, :vw], orc.fr[1]["planes"][k][:, :vw]), "plane %d" % k
    for f in range(4):
        assert np.array_equal(ref.get(f, 0), orc.fr[f]["intraCost"])
        assert np.array_equal(ref.get(f, 1, dtype=np.uint8), orc.fr[f]["intraMode"])
        assert np.array_equal(ref.get(f, 2, 0, 0, np.uint16), orc.fr[f]["lowresCosts"][(0, 0)])
        assert np.array_equal(ref.get(f, 3, 0, 0, np.int32, ref.h8), orc.fr[f]["rowSatds"][(0, 0)])
    for (p0, p1, b) in TRIPLES:
        a, c = ref.cost(p0, p1, b), orc.cost(p0, p1, b)
        assert a == c, ((p0, p1, b), a, c)
        d0, d1 = b - p0, p1 - b
        assert np.array_equal(ref.get(b, 2, d0, d1, np.uint16), orc.fr[b]["lowresCosts"][(d0, d1)]), (p0, p1, b)
        assert np.array_equal(ref.get(b, 3, d0, d1, np.int32, ref.h8), orc.fr[b]["rowSatds"][(d0, d1)])
        assert np.array_equal(ref.get(b, 4, 0, d0, np.int32, 2 * ref.ncu).reshape(-1, 2), orc.fr[b]["mvs"][(0, d0)])
        assert np.array_equal(ref.get(b, 5, 0, d0), orc.fr[b]["mvcosts"][(0, d0)])
        if p1 > b:
            assert np.array_equal(ref.get(b, 4, 1, d1, np.int32, 2 * ref.ncu).reshape(-1, 2), orc.fr[b]["mvs"][(1, d1)])
        st = ref.get(b, 6, d0, d1, np.int64, 3)
        assert (int(st[0]), int(st[1])) == orc.fr[b]["costEst"][(d0, d1)]
        assert int(st[2]) == orc.fr[b]["intraMbs"].get(d0, 0)
