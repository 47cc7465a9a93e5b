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
                ("merange", I), ("rect", I), ("qp", I), ("amp", I)]


class DrvFrame(C.Structure):
    _fields_ = [("p", FsParams), ("fenc", P), ("refs", P * 16), ("field", P), ("mvcost", P),
                ("njobs", I), ("jobs", P), ("me_out", P),
                ("ncu", I), ("cus", P), ("cu_jobs", P), ("cu_coef_off", P), ("coef", P), ("recon", P * 4),
                ("cu_sse", P), ("cu_numsig", P), ("cu_ref", P), ("intra_cost", P),
                ("threads", I), ("next", I), ("stage", I),
                ("chroma", I), ("cstride", I), ("fencC", P * 2), ("refC", (P * 2) * 16),
                ("maxCtuRows", I), ("njobs_run", I), ("ncu_run", I)]


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


def gen_chroma(W, H, n, plane, bits=8, seed=265, noise=False):
    """Cb (plane 1) / Cr (plane 2) of frame n, 4:2:0: smooth sinusoids moving with the clip's global motion (half the luma
    displacement), BASELINE.md generator."""
    cw, ch = W // 2, H // 2
    rng = np.random.default_rng(seed + 1000 * plane + n)
    if noise:
        c = rng.integers(0, 256, (ch, cw))
    else:
        yy, xx = np.mgrid[0:ch, 0:cw]
        ph = 0.7 * plane
        c = 128 + 30 * np.sin((xx + 1.5 * n) / 23.0 + ph) + 25 * np.cos((yy - 1.0 * n) / 19.0 - ph) + rng.integers(-4, 5, (ch, cw))
    c = np.clip(c, 0, 255).astype(np.int64) << (bits - 8)
    return c.astype(pixel_dtype(bits))


def pad_plane_c(img, depth, luma_stride):
    """Margin-extended chroma plane: stride = luma stride / 2, margins = half the luma margins (picyuv.cpp:87-94)."""
    H, W = img.shape
    mx, my, stride = MARGIN_X // 2, MARGIN_Y // 2, luma_stride // 2
    buf = np.zeros((H + 2 * my, stride), pixel_dtype(depth))
    buf[my:my + H, mx:mx + W] = img
    buf[my:my + H, :mx] = img[:, :1]
    buf[my:my + H, mx + W:mx + W + mx] = img[:, -1:]
    buf[:my, :] = buf[my:my + 1, :]
    buf[my + H:, :] = buf[my + H - 1:my + H, :]
    return buf, stride, my * stride + mx


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


def make_field(W, H, numRefs, seed=7):
    """16x16-granular qpel predictor field: global motion (-3*(r+1), 2*(r+1)) px plus jitter."""
    rng = np.random.default_rng(seed)
    fw, fh = (W + 15) // 16, (H + 15) // 16
    f = np.zeros((numRefs, fh, fw, 2), np.int16)
    for r in range(numRefs):
        f[r, :, :, 0] = -12 * (r + 1) + rng.integers(-6, 7, (fh, fw))
        f[r, :, :, 1] = 8 * (r + 1) + rng.integers(-6, 7, (fh, fw))
    return f


class Workload:
    """Inputs of one analysed frame (host side)."""

    def __init__(self, W, H, depth=8, numRefs=4, method=3, subme=3, merange=57, rect=1, qp=30, noise=False, seed=265, chroma=False, amp=0):
        self.W, self.H, self.depth = W, H, depth
        self.chroma = chroma
        self.stride = stride_for(W)
        self.params = dict(width=W, height=H, stride=self.stride, numRefs=numRefs, method=method, subme=subme,
                           merange=merange, rect=rect, qp=qp, amp=amp)
        cur = gen_luma(W, H, numRefs, bits=depth, seed=seed, noise=noise)
        self.fenc, _, self.org = pad_plane(cur, depth)
        self.refs = []
        for r in range(numRefs):
            img = gen_luma(W, H, numRefs - 1 - r, bits=depth, seed=seed, noise=noise)
            self.refs.append(pad_plane(img, depth)[0])
        self.field = make_field(W, H, numRefs)
        if chroma:
            # 4:2:0 planes for the chroma-SATD term of subpelCompare: [Cb, Cr] of the source and of every reference
            self.fencC, self.refC = [], []
            for pl in (1, 2):
                buf, self.cstride, self.corg = pad_plane_c(gen_chroma(W, H, numRefs, pl, bits=depth, seed=seed, noise=noise), depth, self.stride)
                self.fencC.append(buf)
            for r in range(numRefs):
                self.refC.append([pad_plane_c(gen_chroma(W, H, numRefs - 1 - r, pl, bits=depth, seed=seed, noise=noise), depth, self.stride)[0]
                                  for pl in (1, 2)])


def lambda_for(qp, depth):
    # x265_lambda_tab (constants.cpp:33-50): 2^(qp/6-2) * 2^(depth-8), rounded to 4 decimals like the table
    return round(2.0 ** (qp / 6.0 - 2.0) * (1 << (depth - 8)), 4)


def cpu_analyse(lib, fn_name, wl, mvcost_tab, threads=1, stages=7, max_ctu_rows=0):
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
    if getattr(wl, "chroma", False):
        f.chroma, f.cstride = 1, wl.cstride
        for k in range(2):
            f.fencC[k] = wl.fencC[k].ctypes.data + wl.corg * es
            for r in range(len(wl.refs)):
                f.refC[r][k] = wl.refC[r][k].ctypes.data + wl.corg * es
    W, H, nref = wl.W, wl.H, wl.params["numRefs"]
    ctus = ((W + 63) // 64) * ((H + 63) // 64)
    maxjobs = ctus * nref * 593
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
    f.maxCtuRows = max_ctu_rows
    fn = getattr(lib, fn_name)
    fn.argtypes = [C.POINTER(DrvFrame), I]
    fn(C.byref(f), stages)
    nj, nc = f.njobs, f.ncu
    ncoef = int(out["cu_coef_off"][nc - 1] + int(out["cus"][nc - 1][2]) ** 2) if nc else 0
    res = dict(njobs_run=f.njobs_run, ncu_run=f.ncu_run, njobs=nj, ncu=nc, ncoef=ncoef, jobs=out["jobs"][:nj], me_out=out["me_out"][:nj], cus=out["cus"][:nc],
               cu_jobs=out["cu_jobs"][:nc], cu_coef_off=out["cu_coef_off"][:nc], coef=out["coef"][:ncoef],
               recon=out["recon"], cu_sse=out["cu_sse"][:nc], cu_numsig=out["cu_numsig"][:nc], cu_ref=out["cu_ref"][:nc],
               intra_cost=out["intra_cost"][:nc])
    return res
