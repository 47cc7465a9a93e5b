"""Frame-level workload helpers shared by tests and bench.py: synthetic frames (BASELINE.md
generator), margin-extended planes, the predictor field, and the ctypes view of the CPU drivers
(oracle/frame_driver.h instantiated by the oracle and by the reference shim)."""
import ctypes as C
import numpy as np

from common import P, I, ptr, pixel_dtype

MARGIN_X, MARGIN_Y = 96, 80          # picyuv.cpp:87-88 for CTU 64
MVRANGE = 65536


class FsParams(C.Structure):
    _fields_ = [("width", I), ("height", I), ("stride", I), ("numRefs", I), ("method", I), ("subme", I),
                ("merange", I), ("rect", I), ("qp", I)]


class DrvFrame(C.Structure):
    _fields_ = [("p", FsParams), ("fenc", P), ("refs", P * 16), ("field", P), ("mvcost", P),
                ("njobs", I), ("jobs", P), ("me_out", P),
                ("ncu", I), ("cus", P), ("cu_jobs", P), ("cu_coef_off", P), ("coef", P), ("recon", P * 4),
                ("cu_sse", P), ("cu_numsig", P), ("cu_ref", P), ("intra_cost", P),
                ("threads", I), ("next", I), ("stage", I)]


def stride_for(width):
    return (width + 2 * MARGIN_X + 63) // 64 * 64


def gen_luma(W, H, n, s1=37.0, s2=29.0, bits=8, seed=265, noise=False):
    """Luma of frame n of the BASELINE.md synthetic clip (global motion (3,-2) px/frame)."""
    rng = np.random.default_rng(seed + n)
    if noise:
        y = rng.integers(0, 256, (H, W))
    else:
        yy, xx = np.mgrid[0:H, 0:W]
        y = 128 + 60 * np.sin((xx + 3 * n) / s1) + 40 * np.cos((yy - 2 * n) / s2) + rng.integers(-6, 7, (H, W))
    y = np.clip(y, 0, 255).astype(np.int64) << (bits - 8)
    return y.astype(pixel_dtype(bits))


def pad_plane(img, depth):
    """Margin-extended plane (replicated borders, pixel.cpp:1027-1041); returns (buffer, stride, origin element offset)."""
    H, W = img.shape
    stride = stride_for(W)
    buf = np.zeros((H + 2 * MARGIN_Y, stride), pixel_dtype(depth))
    buf[MARGIN_Y:MARGIN_Y + H, MARGIN_X:MARGIN_X + W] = img
    buf[MARGIN_Y:MARGIN_Y + H, :MARGIN_X] = img[:, :1]
    buf[MARGIN_Y:MARGIN_Y + H, MARGIN_X + W:MARGIN_X + W + MARGIN_X] = img[:, -1:]
    buf[:MARGIN_Y, :] = buf[MARGIN_Y:MARGIN_Y + 1, :]
    buf[MARGIN_Y + H:, :] = buf[MARGIN_Y + H - 1:MARGIN_Y + H, :]
    return buf, stride, MARGIN_Y * stride + MARGIN_X


def make_field(W, H, numRefs, seed=7, dist=0):
    """16x16-granular qpel predictor field: (-3*d, 2*d) px plus jitter with d = r + 1 + dist.  dist = 0 is the field of
    the frame right after the newest reference; bench.py passes dist = -k for the frame k positions later so that the
    field keeps the same offset from that frame's true global motion ((+3, -2) px per frame of distance)."""
    rng = np.random.default_rng(seed)
    fw, fh = (W + 15) // 16, (H + 15) // 16
    f = np.zeros((numRefs, fh, fw, 2), np.int16)
    for r in range(numRefs):
        f[r, :, :, 0] = -12 * (r + 1 + dist) + rng.integers(-6, 7, (fh, fw))
        f[r, :, :, 1] = 8 * (r + 1 + dist) + rng.integers(-6, 7, (fh, fw))
    return f


class Workload:
    """Inputs of one analysed frame (host side)."""

    def __init__(self, W, H, depth=8, numRefs=4, method=3, subme=3, merange=57, rect=1, qp=30, noise=False, seed=265):
        self.W, self.H, self.depth = W, H, depth
        self.stride = stride_for(W)
        self.params = dict(width=W, height=H, stride=self.stride, numRefs=numRefs, method=method, subme=subme,
                           merange=merange, rect=rect, qp=qp)
        cur = gen_luma(W, H, numRefs, bits=depth, seed=seed, noise=noise)
        self.fenc, _, self.org = pad_plane(cur, depth)
        self.refs = []
        for r in range(numRefs):
            img = gen_luma(W, H, numRefs - 1 - r, bits=depth, seed=seed, noise=noise)
            self.refs.append(pad_plane(img, depth)[0])
        self.field = make_field(W, H, numRefs)


def lambda_for(qp, depth):
    # x265_lambda_tab (constants.cpp:33-50): 2^(qp/6-2) * 2^(depth-8), rounded to 4 decimals like the table
    return round(2.0 ** (qp / 6.0 - 2.0) * (1 << (depth - 8)), 4)


def cpu_analyse(lib, fn_name, wl, mvcost_tab, threads=1, stages=7):
    """Run the CPU driver (oracle: orc_analyse_frame, reference: x265ref_analyse_frame)."""
    dt = pixel_dtype(wl.depth)
    f = DrvFrame()
    for k, v in wl.params.items():
        setattr(f.p, k, v)
    es = np.dtype(dt).itemsize
    f.fenc = wl.fenc.ctypes.data + wl.org * es
    for r, ref in enumerate(wl.refs):
        f.refs[r] = ref.ctypes.data + wl.org * es
    f.field = wl.field.ctypes.data
    f.mvcost = mvcost_tab.ctypes.data + MVRANGE * 2
    W, H, nref = wl.W, wl.H, wl.params["numRefs"]
    ctus = ((W + 63) // 64) * ((H + 63) // 64)
    maxjobs = ctus * nref * 425
    maxcu = ctus * 85
    from x265_b200.lib import ME_JOB
    out = dict(
        jobs=np.zeros(maxjobs, ME_JOB), me_out=np.zeros((maxjobs, 4), np.int32),
        cus=np.zeros((maxcu, 3), np.int16), cu_jobs=np.zeros((maxcu, nref), np.int32),
        cu_coef_off=np.zeros(maxcu, np.int64), coef=np.zeros(ctus * 4096 * 4, np.int16),
        recon=[np.zeros_like(wl.fenc) for _ in range(4)],
        cu_sse=np.zeros(maxcu, np.uint64), cu_numsig=np.zeros(maxcu, np.uint32), cu_ref=np.zeros(maxcu, np.int32),
        intra_cost=np.zeros((maxcu, 36), np.uint32))
    f.jobs = out["jobs"].ctypes.data; f.me_out = out["me_out"].ctypes.data
    f.cus = out["cus"].ctypes.data; f.cu_jobs = out["cu_jobs"].ctypes.data
    f.cu_coef_off = out["cu_coef_off"].ctypes.data; f.coef = out["coef"].ctypes.data
    for d in range(4):
        f.recon[d] = out["recon"][d].ctypes.data + wl.org * es
    f.cu_sse = out["cu_sse"].ctypes.data; f.cu_numsig = out["cu_numsig"].ctypes.data
    f.cu_ref = out["cu_ref"].ctypes.data; f.intra_cost = out["intra_cost"].ctypes.data
    f.threads = threads
    fn = getattr(lib, fn_name)
    fn.argtypes = [C.POINTER(DrvFrame), I]
    fn(C.byref(f), stages)
    nj, nc = f.njobs, f.ncu
    ncoef = int(out["cu_coef_off"][nc - 1] + int(out["cus"][nc - 1][2]) ** 2) if nc else 0
    res = dict(njobs=nj, ncu=nc, ncoef=ncoef, jobs=out["jobs"][:nj], me_out=out["me_out"][:nj], cus=out["cus"][:nc],
               cu_jobs=out["cu_jobs"][:nc], cu_coef_off=out["cu_coef_off"][:nc], coef=out["coef"][:ncoef],
               recon=out["recon"], cu_sse=out["cu_sse"][:nc], cu_numsig=out["cu_numsig"][:nc], cu_ref=out["cu_ref"][:nc],
               intra_cost=out["intra_cost"][:nc])
    return res
