"""GPU parity of x265cu_pred_cost_batch -- AMVP candidate SADs (Search::selectMVP, search.cpp:1992-2023), merge candidate
SATD + chroma SATD (Search::mergeEstimation, search.cpp:1901-1960; Predict::motionCompensation uni / bi) and the bidir
estimates of predInterSearch (search.cpp:2474-2607) -- against the oracle, which tests/test_oracle_vs_ref.py::test_pred_cost
pins to the real Predict / MotionEstimate classes.  Every PU size, both launch classes (warp per small PU, CTA per large PU),
full / half / quarter-pel vectors, 8- and 10-bit, one list and two, several references."""
import numpy as np
import pytest

import x265_b200
from common import load_oracle, pixel_dtype

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    return x265_b200.load()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("smooth", [True, False])
def test_pred_cost_batch(cu, depth, smooth):
    import pred_helpers as ph
    from x265_b200.lib import PRED_JOB, PRED_CHROMA, PRED_AVG_PP
    O = load_oracle(depth)
    rng = np.random.default_rng(900 + depth + int(smooth))
    pl = ph.make_planes(depth, rng, smooth, nrefs=3)
    jobs, want = [], []
    for (w, h) in ph.LUMA_PUS:
        for kind in ("amvp", "merge_uni", "merge_bi", "bidir_pp"):
            for _ in range(4):
                jb = ph.random_job(pl, rng, w, h, kind)
                jobs.append(jb); want.append(ph.oracle_cost(O, pl, jb))
    n = len(jobs)
    j = np.zeros(n, PRED_JOB)
    for k, jb in enumerate(jobs):
        j[k]["offset"] = pl["org"] + jb["by"] * pl["stride"] + jb["bx"]
        j[k]["pw"], j[k]["ph"] = jb["w"], jb["h"]
        j[k]["ref0"] = jb["ref"][0] if 0 in jb["lists"] else -1
        j[k]["ref1"] = jb["ref"][1] if 1 in jb["lists"] else -1
        j[k]["cost"] = jb["cost"]
        j[k]["flags"] = (PRED_CHROMA if jb["chroma"] else 0) | (PRED_AVG_PP if jb["biAvgPP"] else 0)
        j[k]["mv0"] = jb["mvs"][0]; j[k]["mv1"] = jb["mvs"][1]
    d_fenc = cu.to_device(pl["fenc"])
    d_refs = [cu.to_device(r) for r in pl["refs"]]
    d_fc = [cu.to_device(c) for c in pl["fc"]]
    d_rc = [[cu.to_device(c) for c in rc] for rc in pl["rc"]]
    tab = cu.to_device(np.array([d.ptr for d in d_refs], np.uint64))
    tcb = cu.to_device(np.array([rc[0].ptr for rc in d_rc], np.uint64))
    tcr = cu.to_device(np.array([rc[1].ptr for rc in d_rc], np.uint64))
    d_jobs, d_out = cu.to_device(j), cu.alloc(4 * n)
    cu.pred_cost_batch(depth, d_fenc, pl["stride"], tab, pl["stride"], (d_fc[0], d_fc[1], tcb, tcr, pl["cstride"]), d_jobs, n, d_out)
    got = d_out.download(np.int32)
    bad = np.nonzero(got != np.array(want, np.int32))[0]
    assert bad.size == 0, [(jobs[i]["w"], jobs[i]["h"], jobs[i]["lists"], jobs[i]["cost"], jobs[i]["chroma"], jobs[i]["biAvgPP"], jobs[i]["mvs"], int(got[i]), want[i]) for i in bad[:6]]
    # without chroma planes the chroma term is skipped (the luma-only costs of the same jobs)
    cu.pred_cost_batch(depth, d_fenc, pl["stride"], tab, pl["stride"], None, d_jobs, n, d_out)
    got2 = d_out.download(np.int32)
    luma_only = np.array([ph.oracle_cost(O, pl, dict(jb, chroma=0)) for jb in jobs], np.int32)
    assert np.array_equal(got2, luma_only)
    for d in [d_fenc, tab, tcb, tcr, d_jobs, d_out] + d_refs + d_fc + [c for rc in d_rc for c in rc]:
        d.free()
