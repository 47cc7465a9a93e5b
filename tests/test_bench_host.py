"""Host-side helpers of bench.py that run on both arms: the prediction-cost job list derived from the motion-search jobs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_pred_jobs_of_layout_and_clipping():
    import bench
    from x265_b200.lib import ME_JOB, PRED_JOB, PRED_CHROMA, PRED_AVG_PP
    rng = np.random.default_rng(3)
    n, nrefs = 50, 4
    jobs = np.zeros(n, ME_JOB)
    jobs["offset"] = rng.integers(0, 1 << 20, n); jobs["ref"] = rng.integers(0, nrefs, n)
    jobs["pw"] = 16; jobs["ph"] = 8
    jobs["mvmin"] = rng.integers(-60, -10, (n, 2)); jobs["mvmax"] = rng.integers(10, 60, (n, 2))
    jobs["qmvp"] = rng.integers(-400, 400, (n, 2)); jobs["mvc"] = rng.integers(-400, 400, (n, 8))
    me = np.zeros((n, 4), np.int32); me[:, 1:3] = rng.integers(-200, 200, (n, 2))
    for chroma in (True, False):
        pj = bench.pred_jobs_of(jobs, me, nrefs, chroma)
        assert pj.dtype == PRED_JOB and len(pj) == 4 * n
        q = pj.reshape(n, 4)
        for k in range(4):
            assert np.array_equal(q[:, k]["offset"], jobs["offset"]) and np.array_equal(q[:, k]["ref0"], jobs["ref"])
            assert (q[:, k]["pw"] == 16).all() and (q[:, k]["ph"] == 8).all()
        lo, hi = jobs["mvmin"].astype(np.int32) * 4, jobs["mvmax"].astype(np.int32) * 4
        for k, src in ((0, jobs["qmvp"]), (1, jobs["mvc"][:, 0:2]), (2, jobs["mvc"][:, 2:4])):
            assert np.array_equal(q[:, k]["mv0"], np.clip(src, lo, hi))            # AMVP / merge vectors stay inside the search window
            assert (q[:, k]["ref1"] == -1).all()
        assert (q[:, 0]["cost"] == 0).all() and (q[:, 1]["cost"] == 0).all() and (q[:, 2]["cost"] == 1).all() and (q[:, 3]["cost"] == 1).all()
        assert np.array_equal(q[:, 3]["mv0"], me[:, 1:3]) and np.array_equal(q[:, 3]["ref1"], (jobs["ref"] + 1) % nrefs)
        assert (q[:, 2]["flags"] == (PRED_CHROMA if chroma else 0)).all()
        assert (q[:, 3]["flags"] == (PRED_CHROMA if chroma else PRED_AVG_PP)).all()
    c = bench.pred_checks(np.arange(10))
    assert c["jobs"] == 10 and c["cost_sum"] == 45
