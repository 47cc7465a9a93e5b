"""GPU parity of the frame-level CTU-analysis pipeline (x265cu_analyser_*: job build -> batched
motionEstimate -> fused MC/residual/DCT/quant/IDCT/recon -> intra search) against the oracle driver,
which test_frame_oracle_vs_ref.py pins to the real reference code.  Every output array is compared
bit-exactly; full-size (2160p) runs are checked through size-independent properties in bench.py."""
import numpy as np
import pytest

from common import load_oracle
from frame_helpers import Workload, cpu_analyse, lambda_for, MARGIN_X, MARGIN_Y
from me_helpers import mvcost_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    import x265_b200
    return x265_b200.load()


def run_gpu(cu, wl, qp):
    import x265_b200
    p = wl.params
    an = x265_b200.Analyser(cu, wl.W, wl.H, depth=wl.depth, numRefs=p["numRefs"], method=p["method"], subme=p["subme"],
                            merange=p["merange"], rect=p["rect"], qp=qp, lam=lambda_for(qp, wl.depth))
    H, W = wl.H, wl.W
    for r, ref in enumerate(wl.refs):
        an.set_ref(r, ref[MARGIN_Y:MARGIN_Y + H, MARGIN_X:MARGIN_X + W])
    fenc = np.ascontiguousarray(wl.fenc[MARGIN_Y:MARGIN_Y + H, MARGIN_X:MARGIN_X + W])
    an.analyse(fenc, wl.field)
    return an


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("noise", [False, True])
@pytest.mark.parametrize("size", [(200, 136), (256, 128)])
def test_frame_analysis(cu, depth, noise, size):
    O = load_oracle(depth)
    qp = 30
    W, H = size
    wl = Workload(W, H, depth=depth, numRefs=2, method=3, subme=3, merange=57, rect=1, qp=qp, noise=noise)
    tab = mvcost_table(O, lambda_for(qp, depth))
    want = cpu_analyse(O, "orc_analyse_frame", wl, tab, threads=8)
    an = run_gpu(cu, wl, qp)
    assert an.njobs == want["njobs"] and an.ncu == want["ncu"] and an.ncoef == want["ncoef"] and an.stride == wl.stride
    jobs = an.fetch("jobs")
    assert np.array_equal(jobs, want["jobs"]), "job lists differ"
    assert np.array_equal(an.fetch("cu_jobs").reshape(-1, 2), want["cu_jobs"])
    # border extension of the uploaded source plane
    vw = W + 2 * MARGIN_X      # columns past the right margin are stride padding
    assert np.array_equal(an.fetch("fenc").reshape(wl.fenc.shape)[:, :vw], wl.fenc[:, :vw])
    me = an.fetch("me_out").reshape(-1, 4)
    bad = np.nonzero((me[:, :3] != want["me_out"][:, :3]).any(axis=1))[0]
    assert len(bad) == 0, (len(bad), jobs[bad[:3]].tolist(), me[bad[:3]].tolist(), want["me_out"][bad[:3]].tolist())
    # packed D2H form
    assert np.array_equal(an.me_packed[:, 0], want["me_out"][:, 0])
    assert np.array_equal(an.me_packed[:, 1].view(np.uint32) & 0xffff, want["me_out"][:, 1].astype(np.uint32) & 0xffff)
    assert np.array_equal(an.me_packed[:, 1].view(np.uint32) >> 16, want["me_out"][:, 2].astype(np.uint32) & 0xffff)
    assert np.array_equal(an.cu_ref, want["cu_ref"])
    assert np.array_equal(an.fetch("coef"), want["coef"])
    assert np.array_equal(an.cu_numsig, want["cu_numsig"])
    assert np.array_equal(an.cu_sse, want["cu_sse"])
    for d in range(4):
        got = an.fetch("recon%d" % d).reshape(wl.fenc.shape)[MARGIN_Y:MARGIN_Y + H, MARGIN_X:MARGIN_X + W]
        exp = want["recon"][d][MARGIN_Y:MARGIN_Y + H, MARGIN_X:MARGIN_X + W]
        assert np.array_equal(got, exp), "recon depth %d" % d
    assert np.array_equal(an.intra_cost, want["intra_cost"])
    an.close()
