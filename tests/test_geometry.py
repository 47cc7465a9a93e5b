"""CPU check of the frame analyser's HOST side: the geometry tables x265cu_analyser_create builds (x265_b200/csrc/geometry.h,
compiled here with g++ through tests/geometry_capi.cpp) against the oracle's independent enumeration (oracle/frame_spec.h via
orc_analyse_frame's prepare step), for every BASELINE picture size, rect / AMP on and off; plus the layout of the ABI structs
the ctypes binding mirrors.  The GPU tests compare the same lists on the device; this one needs no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from common import load_oracle
from frame_helpers import Workload, cpu_analyse, lambda_for, stride_for
from me_helpers import mvcost_table

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def geo(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("geo") / "libgeo.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "geometry_capi.cpp")])
    return C.CDLL(so)


def build(geo, W, H, nref, rect, amp):
    counts = np.zeros(5, np.int64)
    geo.geo_build(W, H, stride_for(W), nref, rect, amp, counts.ctypes.data_as(C.c_void_p))
    nj, ncu, ntu, rows, ncoef = [int(x) for x in counts]

    def get(what, dtype, n):
        out = np.zeros(n, dtype)
        assert geo.geo_get(what, out.ctypes.data_as(C.c_void_p)) == 0
        return out
    return dict(njobs=nj, ncu=ncu, ntu=ntu, rows=rows, ncoef=ncoef,
                pus=get(0, np.int32, nj * 6).reshape(-1, 6), cus=get(1, np.int64, ncu * 4).reshape(-1, 4),
                tus=get(2, np.int32, ntu * 3).reshape(-1, 3), cu_jobs=get(3, np.int32, ncu * nref).reshape(-1, nref),
                rowJob=get(4, np.int32, rows + 1), rowCu=get(5, np.int32, rows + 1), rowTu=get(6, np.int32, rows + 1))


@pytest.mark.parametrize("size", [(200, 136), (352, 288), (1920, 1080), (3840, 2160)])
@pytest.mark.parametrize("rect,amp", [(1, 0), (0, 0), (1, 1), (0, 1)])
def test_geometry_matches_oracle(geo, size, rect, amp):
    W, H = size
    nref = 2 if W > 1920 else 3
    O = load_oracle(8)
    wl = Workload(W, H, depth=8, numRefs=nref, method=3, subme=3, merange=57, rect=rect, qp=30, amp=amp)
    want = cpu_analyse(O, "orc_analyse_frame", wl, mvcost_table(O, lambda_for(30, 8)), threads=1, stages=0)     # prepare only
    g = build(geo, W, H, nref, rect, amp)
    assert (g["njobs"], g["ncu"], g["ncoef"]) == (want["njobs"], want["ncu"], want["ncoef"])
    j = want["jobs"]
    assert np.array_equal(g["pus"][:, 0], j["offset"]) and np.array_equal(g["pus"][:, 5], j["ref"])
    assert np.array_equal(g["pus"][:, 3], j["pw"]) and np.array_equal(g["pus"][:, 4], j["ph"])
    assert np.array_equal(g["cus"][:, :3], want["cus"].astype(np.int64))
    assert np.array_equal(g["cus"][:, 3], want["cu_coef_off"])
    assert np.array_equal(g["cu_jobs"], want["cu_jobs"])
    # every PU lies inside its CU, every CU job is the CU's 2Nx2N PU
    st = stride_for(W)
    px, py = g["pus"][:, 0] % st, g["pus"][:, 0] // st
    assert (px >= g["pus"][:, 1]).all() and (py >= g["pus"][:, 2]).all()
    k = g["cu_jobs"][:, 0]
    assert np.array_equal(g["pus"][k, 3], g["cus"][:, 2]) and np.array_equal(g["pus"][k, 4], g["cus"][:, 2])
    # TUs: 32x32 tiles of 64x64 CUs, else the CU itself; coefficient space = sum of CU areas
    assert g["ntu"] == int((np.maximum(g["cus"][:, 2] // 32, 1) ** 2).sum())
    assert g["ncoef"] == int((g["cus"][:, 2] ** 2).sum())
    # CTU-row prefix tables: monotone, end at the totals, slices cover whole CTU rows
    assert g["rows"] == (H + 63) // 64
    for key, total in (("rowJob", g["njobs"]), ("rowCu", g["ncu"]), ("rowTu", g["ntu"])):
        assert g[key][0] == 0 and g[key][-1] == total and (np.diff(g[key]) >= 0).all()
    for r in range(g["rows"]):
        rows_of_cus = g["cus"][g["rowCu"][r]:g["rowCu"][r + 1], 1] // 64
        assert (rows_of_cus == r).all()
        rows_of_pus = (g["pus"][g["rowJob"][r]:g["rowJob"][r + 1], 0] // st) // 64
        assert (rows_of_pus == r).all()


@pytest.mark.parametrize("size", [(200, 136), (352, 288), (1920, 1080), (3840, 2160)])
@pytest.mark.parametrize("rect,amp", [(1, 0), (1, 1), (0, 0)])
def test_window_groups_partition_the_jobs(geo, size, rect, amp):
    """Job groups of the shared-memory-window integer search (me_window.cuh): every PU job is in exactly one group, a group's
    jobs share the reference and lie inside one 64x64 CU (class 0), one 32x32 CU (class 1) or one 16x16 cell (class 2), are ordered largest
    first, and the per-CTU-row group ranges cover exactly the jobs of those rows (what a row shard launches)."""
    W, H = size
    nref = 2
    g = build(geo, W, H, nref, rect, amp)
    st = stride_for(W)
    seen = np.zeros(g["njobs"], np.int32)
    pus = g["pus"]
    px, py = pus[:, 0] % st, pus[:, 0] // st
    for k in range(3):
        cnt = np.zeros(2, np.int32); geo.geo_get(10 + 4 * k, cnt.ctypes.data_as(C.c_void_p))
        ng, nj = int(cnt[0]), int(cnt[1])
        fc = np.zeros(2 * ng, np.int32); geo.geo_get(11 + 4 * k, fc.ctypes.data_as(C.c_void_p)); fc = fc.reshape(-1, 2)
        jobs = np.zeros(nj, np.int32); geo.geo_get(12 + 4 * k, jobs.ctypes.data_as(C.c_void_p))
        rowg = np.zeros(g["rows"] + 1, np.int32); geo.geo_get(13 + 4 * k, rowg.ctypes.data_as(C.c_void_p))
        assert rowg[0] == 0 and rowg[-1] == ng and (np.diff(rowg) >= 0).all()
        np.add.at(seen, jobs, 1)
        assert fc[:, 1].max() <= 48 and fc[:, 1].min() >= 1          # MEW_MAX_GROUP
        for gi in range(0, ng, max(1, ng // 400)):                    # a sample of groups in detail
            f, c = fc[gi]
            ids = jobs[f:f + c]
            assert len(set(pus[ids, 5])) == 1
            x0, y0 = px[ids].min(), py[ids].min()
            x1, y1 = (px[ids] + pus[ids, 3]).max(), (py[ids] + pus[ids, 4]).max()
            lim = (64, 32, 16)[k]
            assert x1 - x0 <= lim and y1 - y0 <= lim and x0 // lim == (x1 - 1) // lim and y0 // lim == (y1 - 1) // lim
            area = pus[ids, 3].astype(np.int64) * pus[ids, 4]
            assert (np.diff(area) <= 0).all()
        for r in range(g["rows"]):
            ids = jobs[fc[rowg[r], 0]:(fc[rowg[r + 1] - 1, 0] + fc[rowg[r + 1] - 1, 1])] if rowg[r + 1] > rowg[r] else jobs[:0]
            assert ((py[ids] // 64) == r).all()
    assert (seen == 1).all()
    # shape-sorted job order: a permutation of every CTU row's jobs, PU area non-increasing inside a row
    order = np.zeros(g["njobs"], np.int32); geo.geo_get(22, order.ctypes.data_as(C.c_void_p))
    for r in range(g["rows"]):
        j0, j1 = g["rowJob"][r], g["rowJob"][r + 1]
        seg = order[j0:j1]
        assert np.array_equal(np.sort(seg), np.arange(j0, j1))
        area = pus[seg, 3].astype(np.int64) * pus[seg, 4]
        assert (np.diff(area) <= 0).all()


def test_geometry_8k_ranges(geo):
    """BASELINE configs[4] (7680x4320, 5 references, rect + AMP): counts by formula and every field inside its integer type."""
    W, H, nref = 7680, 4320, 5
    g = build(geo, W, H, nref, 1, 1)
    ctus_full = (W // 64) * (H // 64)                       # 120 x 67 full CTUs + a 32-pixel-high last row
    assert g["rows"] == 68
    per_ctu = 85 * 5 + 21 * 8
    assert g["njobs"] > ctus_full * per_ctu * nref and g["njobs"] < 8160 * per_ctu * nref
    assert g["pus"][:, 0].max() < 2 ** 31 - 1 and g["pus"][:, 0].min() >= 0
    assert g["pus"][:, 1].max() < 32768 and g["cus"][:, 0].max() < 32768
    assert g["ncoef"] == int((g["cus"][:, 2] ** 2).sum()) and g["ncoef"] > 4 * W * (H - 64)
    sizes = {(int(w), int(h)) for w, h in zip(g["pus"][:, 3], g["pus"][:, 4])}
    assert {(64, 16), (64, 48), (16, 64), (48, 64), (32, 8), (32, 24), (8, 32), (24, 32), (16, 4), (16, 12), (4, 16), (12, 16)} <= sizes


def test_abi_struct_layout(geo):
    from x265_b200.lib import AnalysisParams, AnalysisOut, MeChroma, ME_JOB
    out = np.zeros(16, np.int32)
    n = geo.geo_abi_layout(out.ctypes.data_as(C.c_void_p))
    assert n == 11
    assert C.sizeof(AnalysisParams) == out[0] and AnalysisParams.qp.offset == out[1]
    assert AnalysisParams.lam.offset == out[2] and AnalysisParams.amp.offset == out[3]
    assert ME_JOB.itemsize == out[4] == 40
    assert C.sizeof(MeChroma) == out[5] and MeChroma.cstride.offset == out[6]
    assert C.sizeof(AnalysisOut) == out[7]
    assert (out[8], out[9], out[10]) == (12, 16, 8)
