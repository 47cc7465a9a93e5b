"""GPU parity of the lookahead path (BASELINE configs[1]): x265cu_frame_init_lowres ->
x265cu_lowres_intra_batch -> x265cu_lookahead_cost_batch against the oracle restatement, which
test_lookahead_oracle_vs_ref.py pins to the real Lowres / LookaheadTLD / CostEstimateGroup classes."""
import numpy as np
import pytest

from common import load_oracle, pixel_dtype
from frame_helpers import gen_luma, MARGIN_X, MARGIN_Y, MVRANGE
from lookahead_helpers import OracleLookahead, full_plane, lowres_geometry, LOOKAHEAD_LAMBDA

pytestmark = pytest.mark.gpu
TRIPLES = [(0, 1, 1), (0, 2, 1), (0, 2, 2), (0, 3, 1), (0, 3, 2), (0, 3, 3), (1, 3, 2), (2, 3, 3)]


@pytest.fixture(scope="module")
def cu():
    import x265_b200
    return x265_b200.load()


class GpuLookahead:
    """Host-side bookkeeping (which (list, dist) motion fields exist) + device buffers."""

    def __init__(self, cu, frames, depth):
        from x265_b200.lib import LA_INTRA_JOB
        self.cu, self.depth = cu, depth
        H, W = frames[0].shape
        self.w8, self.h8, self.ls = lowres_geometry(W, H)
        self.ncu = self.w8 * self.h8
        es = np.dtype(pixel_dtype(depth)).itemsize
        lw, lh = self.w8 * 8, self.h8 * 8
        rows = lh + 2 * MARGIN_Y
        self.lorg = (MARGIN_Y * self.ls + MARGIN_X) * es
        self.fr = []
        jobs = np.zeros(len(frames), LA_INTRA_JOB)
        for i, img in enumerate(frames):
            full, fs, forg = full_plane(img, depth)
            d_full = cu.to_device(full)
            planes = [cu.alloc(self.ls * rows * es) for _ in range(4)]
            for p in planes:
                cu.check(cu.L.x265cu_memset(cu.ctx, p.ptr, 0, p.nbytes))
            cu.check(cu.L.x265cu_frame_init_lowres(cu.ctx, depth, d_full.ptr + forg * es, fs, *[p.ptr + self.lorg for p in planes],
                                                   self.ls, lw, lh, MARGIN_X, MARGIN_Y))
            f = dict(planes=planes, intraCost=cu.alloc(4 * self.ncu), intraMode=cu.alloc(self.ncu), lc0=cu.alloc(2 * self.ncu),
                     rs0=cu.alloc(4 * self.h8), out0=cu.alloc(16), mvs={}, mvcosts={}, res={})
            j = jobs[i]
            j["plane0"] = planes[0].ptr + self.lorg; j["invQscale"] = 0; j["intraCost"] = f["intraCost"].ptr; j["intraMode"] = f["intraMode"].ptr
            j["lowresCosts"] = f["lc0"].ptr; j["rowSatds"] = f["rs0"].ptr; j["out"] = f["out0"].ptr
            self.fr.append(f)
            cu.sync(); d_full.free()
        d_jobs = cu.to_device(jobs)
        cu.check(cu.L.x265cu_lowres_intra_batch(cu.ctx, depth, d_jobs.ptr, len(frames), self.ls, self.w8, self.h8, int(LOOKAHEAD_LAMBDA[depth])))
        cu.sync()
        self.tab = cu.to_device(cu.mvcost_table(LOOKAHEAD_LAMBDA[depth], MVRANGE))

    def cost(self, p0, p1, b):
        return self.cost_batch([(p0, p1, b)])[0]

    def cost_batch(self, triples):
        """One launch for all triples (they must not share a motion field that still has to be searched)."""
        from x265_b200.lib import LA_JOB
        cu = self.cu
        todo = [t for t in triples if (t[2] - t[0], t[1] - t[2]) not in self.fr[t[2]]["res"]]
        jobs = np.zeros(max(len(todo), 1), LA_JOB)
        bufs = []
        for n, (p0, p1, b) in enumerate(todo):
            f = self.fr[b]
            d0, d1 = b - p0, p1 - b
            j = jobs[n]
            for k in range(4):
                j["fenc"][k] = f["planes"][k].ptr + self.lorg
                j["ref0"][k] = self.fr[p0]["planes"][k].ptr + self.lorg
                j["ref1"][k] = self.fr[p1]["planes"][k].ptr + self.lorg
            j["bidir"] = int(b < p1)
            for lst, dist in ((0, d0), (1, d1)):
                new = (lst, dist) not in f["mvs"]
                if new:
                    f["mvs"][(lst, dist)] = cu.alloc(8 * self.ncu); f["mvcosts"][(lst, dist)] = cu.alloc(4 * self.ncu)
                    cu.check(cu.L.x265cu_memset(cu.ctx, f["mvs"][(lst, dist)].ptr, 0, 8 * self.ncu))
                    cu.check(cu.L.x265cu_memset(cu.ctx, f["mvcosts"][(lst, dist)].ptr, 0, 4 * self.ncu))
                j["doSearch%d" % lst] = int(new and (lst == 0 or p1 > b))
                j["mvs"][lst] = f["mvs"][(lst, dist)].ptr; j["mvcosts"][lst] = f["mvcosts"][(lst, dist)].ptr
            lc, rs, out = cu.alloc(2 * self.ncu), cu.alloc(4 * self.h8), cu.alloc(24)
            bufs.append((lc, rs, out))
            j["intraCost"] = f["intraCost"].ptr; j["invQscale"] = 0; j["lowresCosts"] = lc.ptr; j["rowSatds"] = rs.ptr; j["out"] = out.ptr
        if todo:
            d_j = cu.to_device(jobs)
            cu.check(cu.L.x265cu_lookahead_cost_batch(cu.ctx, self.depth, d_j.ptr, len(todo), self.ls, self.w8, self.h8, self.tab.ptr + 2 * MVRANGE))
            for (p0, p1, b), (lc, rs, out) in zip(todo, bufs):
                o = out.download(np.int64)
                score = int(o[0])
                if b != p1:
                    score = score * 100 // 130
                self.fr[b]["res"][(b - p0, p1 - b)] = dict(score=score, costEstAq=int(o[1]), intraMbs=int(o[2]), lowresCosts=lc.download(np.uint16),
                                                           rowSatds=rs.download(np.int32))
        return [self.fr[b]["res"][(b - p0, p1 - b)]["score"] for (p0, p1, b) in triples]


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("size,noise", [((416, 240), False), ((200, 136), True), ((960, 544), False)])
def test_lookahead_gpu(cu, depth, size, noise):
    O = load_oracle(depth)
    W, H = size
    frames = [gen_luma(W, H, i, s1=17.0, s2=11.0, bits=depth, noise=noise) for i in range(4)]
    orc = OracleLookahead(O, frames, depth)
    gpu = GpuLookahead(cu, frames, depth)
    vw = orc.w8 * 8 + 2 * MARGIN_X
    for k in range(4):
        got = gpu.fr[1]["planes"][k].download(pixel_dtype(depth)).reshape(orc.fr[1]["planes"][k].shape)
        assert np.array_equal(got[:, :vw], orc.fr[1]["planes"][k][:, :vw]), "lowres plane %d" % k
    for i in range(4):
        assert np.array_equal(gpu.fr[i]["intraCost"].download(np.int32), orc.fr[i]["intraCost"])
        assert np.array_equal(gpu.fr[i]["intraMode"].download(np.uint8), orc.fr[i]["intraMode"])
        assert np.array_equal(gpu.fr[i]["lc0"].download(np.uint16), orc.fr[i]["lowresCosts"][(0, 0)])
        assert np.array_equal(gpu.fr[i]["rs0"].download(np.int32), orc.fr[i]["rowSatds"][(0, 0)])
        assert tuple(int(x) for x in gpu.fr[i]["out0"].download(np.int64)) == orc.fr[i]["costEst"][(0, 0)]
    for (p0, p1, b) in TRIPLES:
        a, c = gpu.cost(p0, p1, b), orc.cost(p0, p1, b)
        d0, d1 = b - p0, p1 - b
        r = gpu.fr[b]["res"][(d0, d1)]
        assert np.array_equal(gpu.fr[b]["mvs"][(0, d0)].download(np.int32).reshape(-1, 2), orc.fr[b]["mvs"][(0, d0)]), (p0, p1, b)
        assert np.array_equal(gpu.fr[b]["mvcosts"][(0, d0)].download(np.int32), orc.fr[b]["mvcosts"][(0, d0)])
        if p1 > b:
            assert np.array_equal(gpu.fr[b]["mvs"][(1, d1)].download(np.int32).reshape(-1, 2), orc.fr[b]["mvs"][(1, d1)])
        assert np.array_equal(r["lowresCosts"], orc.fr[b]["lowresCosts"][(d0, d1)]), (p0, p1, b)
        assert np.array_equal(r["rowSatds"], orc.fr[b]["rowSatds"][(d0, d1)])
        assert a == c and r["costEstAq"] == orc.fr[b]["costEst"][(d0, d1)][1], ((p0, p1, b), a, c)
