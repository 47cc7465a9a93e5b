"""GPU parity of the BATCHED entry points with many jobs per launch (include/x265_b200.h section 2) against
the oracle: the per-call table tests drive the same kernels with one job, these drive the job indexing, the
mixed-size job lists and the vectorised fast paths.  Bit-exact, 8 and 10 bit."""
import ctypes as C
import numpy as np
import pytest

from common import load_oracle, pixel_dtype, ptr, IP, LUMA_PU

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    import x265_b200
    return x265_b200.load()


def _interp_oracle(O, op):
    return getattr(O, "orc_interp_" + op)


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("op", ["hpp", "hps", "vpp", "vps", "vsp", "vss", "hvpp"])
def test_interp_batch_luma(cu, depth, op):
    """A job list mixing PU sizes, fractions, source phases and (for hps) rowExt, all in one launch."""
    from x265_b200.lib import INTERP_JOB
    O = load_oracle(depth)
    rng = np.random.default_rng(11 + depth)
    dt = pixel_dtype(depth)
    SW, SH = 512, 400
    short_src = op in ("vsp", "vss")
    short_dst = op in ("hps", "vps", "vss")
    src = (rng.integers(-(1 << 13), 1 << 13, (SH, SW)).astype(np.int16) if short_src
           else rng.integers(0, 1 << depth, (SH, SW)).astype(dt))
    DW = 640
    jobs, places = [], []
    y = 0
    sizes = [LUMA_PU[i] for i in (1, 2, 3, 4, 7, 8, 9, 10, 11, 12, 19, 20, 23, 24, 0, 5, 6, 13, 14)]
    for k, (w, h) in enumerate(sizes * 3):
        ext = 1 if (op == "hps" and k % 3 == 1) else 0
        rows = h + (7 if ext else 0)
        sx, sy = 16 + int(rng.integers(0, SW - 96)), 8 + int(rng.integers(0, SH - 88))
        dx = int(rng.integers(0, 4)) * (4 if k % 2 else 8)
        ix = 1 + k % 3
        iy = 1 + (k // 3) % 3
        if op != "hvpp" and k % 7 == 0:
            ix = 0
        jobs.append((sy * SW + sx, y * DW + dx, SW, DW, w, h, ix, iy, ext, 8))
        places.append((y, dx, rows, w))
        y += rows + 1
    DH = y + 2
    j = np.zeros(len(jobs), INTERP_JOB)
    for i, t in enumerate(jobs):
        j[i] = t
    ddt = np.int16 if short_dst else dt
    dst0 = rng.integers(0, 100, (DH, DW)).astype(ddt)
    want = dst0.copy()
    fn = _interp_oracle(O, op)
    for (s_off, d_off, ss, ds, w, h, ix, iy, ext, nt) in jobs:
        ps, pd = ptr(src, s_off), ptr(want, d_off)
        if op == "hvpp":
            fn(ps, IP(ss), pd, IP(ds), ix, iy, nt, w, h)
        elif op == "hps":
            fn(ps, IP(ss), pd, IP(ds), ix, ext, nt, w, h)
        else:
            fn(ps, IP(ss), pd, IP(ds), ix, nt, w, h)
    dS, dD, dJ = cu.to_device(src), cu.to_device(dst0), cu.to_device(j)
    cu.interp_batch(depth, op, dS, dD, dJ, len(jobs))
    got = dD.download(ddt).reshape(DH, DW)
    np.testing.assert_array_equal(got, want)
    for d in (dS, dD, dJ):
        d.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_pixelcmp_batch_small_groups(cu, depth):
    """Lists of SMALL blocks: the job-list kernel packs 32 / width jobs per warp pass (width = pow2ceil(work units of the
    largest job among 32 consecutive ones)); uniform 4x4 / 8x8 / 16x16 lists, mixed small lists, n not a multiple of 32,
    sa8d lists mixing sub-8 blocks (SATD path) with 8x8 / 16x16 / 24x32 ones, and the one-plane / int16 costs."""
    from x265_b200.lib import CMP_JOB
    O = load_oracle(depth)
    rng = np.random.default_rng(50 + depth)
    dt = pixel_dtype(depth)
    a = rng.integers(0, 1 << depth, (200, 320)).astype(dt)
    b = rng.integers(0, 1 << depth, (200, 288)).astype(dt)
    sa16 = rng.integers(-600, 600, (200, 320)).astype(np.int16)
    sb16 = rng.integers(-600, 600, (200, 288)).astype(np.int16)
    sse_t = C.c_uint32 if depth == 8 else C.c_uint64
    two = {"sad": (O.orc_sad, C.c_int), "satd": (O.orc_satd, C.c_int), "sa8d": (O.orc_sa8d, C.c_int), "sse_pp": (O.orc_sse_pp, sse_t)}
    lists = {"sad": [[(8, 8)], [(4, 4)], [(16, 16)], [(8, 4), (4, 8), (8, 8), (4, 16)], [(12, 16), (16, 12), (8, 8)]],
             "satd": [[(8, 8)], [(4, 4)], [(16, 16)], [(8, 4), (4, 8), (8, 8), (16, 4)], [(12, 16), (16, 8), (8, 16)]],
             "sa8d": [[(8, 8)], [(16, 16)], [(4, 4), (8, 8), (16, 16), (8, 4)], [(24, 32), (8, 16), (16, 8), (32, 8)], [(4, 8), (8, 4)]],
             "sse_pp": [[(8, 8)], [(4, 4)], [(16, 16)], [(8, 4), (4, 8), (16, 8)]]}
    dA, dB = cu.to_device(a), cu.to_device(b)
    for op, (ofn, res) in two.items():
        ofn.restype = res
        for sizes in lists[op]:
            n = 32 * 3 + 7
            j = np.zeros(n, CMP_JOB); want = np.zeros(n, np.uint64)
            for k in range(n):
                w, h = sizes[int(rng.integers(0, len(sizes)))]
                ao = int(rng.integers(0, 200 - 32)) * 320 + int(rng.integers(0, 320 - 32))
                bo = int(rng.integers(0, 200 - 32)) * 288 + int(rng.integers(0, 288 - 32))
                j[k] = (ao, bo, 320, 288, w, h, 0)
                want[k] = ofn(ptr(a, ao), IP(320), ptr(b, bo), IP(288), w, h)
            dJ, dO = cu.to_device(j), cu.alloc(8 * n)
            cu.pixelcmp_batch(depth, op, dA, dB, dJ, n, dO)
            np.testing.assert_array_equal(dO.download(np.uint64), want, err_msg="%s %s" % (op, sizes))
            dJ.free(); dO.free()
    # one-plane and int16 costs on square blocks (cu[] entries): var, ssd_s, sse_ss, psy
    O.orc_var.restype = C.c_uint64; O.orc_ssd_s.restype = sse_t; O.orc_sse_ss.restype = sse_t; O.orc_psy_cost_pp.restype = C.c_int
    dSA, dSB = cu.to_device(sa16), cu.to_device(sb16)
    for op in ("var", "ssd_s", "sse_ss", "psy"):
        for sizes in ([8], [16], [4, 8, 16] if op != "var" else [8, 16], [32, 8]):
            n = 32 * 2 + 5
            j = np.zeros(n, CMP_JOB); want = np.zeros(n, np.uint64)
            for k in range(n):
                s = sizes[int(rng.integers(0, len(sizes)))]
                ao = int(rng.integers(0, 200 - 32)) * 320 + int(rng.integers(0, 320 - 32))
                bo = int(rng.integers(0, 200 - 32)) * 288 + int(rng.integers(0, 288 - 32))
                j[k] = (ao, bo, 320, 288, s, s, 0)
                if op == "var":      want[k] = O.orc_var(ptr(a, ao), IP(320), s)
                elif op == "ssd_s":  want[k] = O.orc_ssd_s(ptr(sa16, ao), IP(320), s)
                elif op == "sse_ss": want[k] = O.orc_sse_ss(ptr(sa16, ao), IP(320), ptr(sb16, bo), IP(288), s, s)
                else:                want[k] = O.orc_psy_cost_pp(ptr(a, ao), IP(320), ptr(b, bo), IP(288), s)
            dJ, dO = cu.to_device(j), cu.alloc(8 * n)
            pa, pb = (dSA, dSB) if op in ("ssd_s", "sse_ss") else (dA, dB)
            cu.pixelcmp_batch(depth, op, pa, pb, dJ, n, dO)
            np.testing.assert_array_equal(dO.download(np.uint64), want, err_msg="%s %s" % (op, sizes))
            dJ.free(); dO.free()
    for d in (dA, dB, dSA, dSB):
        d.free()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("N", [4, 8, 16, 32])
@pytest.mark.parametrize("layout", ["contiguous", "strided"])
def test_transform_batch(cu, depth, N, layout):
    O = load_oracle(depth)
    rng = np.random.default_rng(N + depth)
    n = 333
    mx = (1 << depth) - 1
    if layout == "contiguous":
        stride, pitch = N, N * N
    else:
        stride, pitch = N + 6, (N + 6) * N + 10
    src = np.zeros(pitch * n + 64, np.int16)
    resid = (rng.integers(0, mx + 1, (n, N, N)) - rng.integers(0, mx + 1, (n, N, N))).astype(np.int16)
    resid[0] = mx; resid[1] = -mx
    for t in range(n):
        for r in range(N):
            src[t * pitch + r * stride: t * pitch + r * stride + N] = resid[t, r]
    want = np.zeros((n, N * N), np.int16)
    for t in range(n):
        O.orc_dct(ptr(src, t * pitch), ptr(want, t * N * N), IP(stride), N)
    dS, dD = cu.to_device(src), cu.alloc(2 * n * N * N)
    cu.transform_batch(depth, "dct", N, dS, dD, stride, pitch, n)
    coef = dD.download(np.int16).reshape(n, N * N)
    np.testing.assert_array_equal(coef, want)
    # inverse: contiguous coefficients in, strided residual out
    wanti = np.zeros(pitch * n + 64, np.int16)
    for t in range(n):
        O.orc_idct(ptr(want, t * N * N), ptr(wanti, t * pitch), IP(stride), N)
    dI = cu.to_device(np.zeros(pitch * n + 64, np.int16))
    cu.transform_batch(depth, "idct", N, dD, dI, stride, pitch, n)
    np.testing.assert_array_equal(dI.download(np.int16), wanti)
    for d in (dS, dD, dI):
        d.free()


@pytest.mark.parametrize("depth", [8, 10])
def test_pixelcmp_batch_mixed(cu, depth):
    from x265_b200.lib import CMP_JOB
    O = load_oracle(depth)
    rng = np.random.default_rng(5 + depth)
    dt = pixel_dtype(depth)
    a = rng.integers(0, 1 << depth, (300, 448)).astype(dt)
    b = rng.integers(0, 1 << depth, (300, 384)).astype(dt)
    for op, ofn, res in (("sad", O.orc_sad, C.c_int), ("satd", O.orc_satd, C.c_int), ("sa8d", O.orc_sa8d, C.c_int),
                         ("sse_pp", O.orc_sse_pp, C.c_uint32 if depth == 8 else C.c_uint64)):
        sizes = LUMA_PU if op != "sa8d" else [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64)]
        j = np.zeros(4 * len(sizes), CMP_JOB)
        want = np.zeros(j.size, np.uint64)
        ofn.restype = res
        for k in range(j.size):
            w, h = sizes[k % len(sizes)]
            ao = int(rng.integers(0, 300 - 64)) * 448 + int(rng.integers(0, 448 - 64))
            bo = int(rng.integers(0, 300 - 64)) * 384 + int(rng.integers(0, 384 - 64))
            j[k] = (ao, bo, 448, 384, w, h, 0)
            want[k] = ofn(ptr(a, ao), IP(448), ptr(b, bo), IP(384), w, h)
        dA, dB, dJ, dO = cu.to_device(a), cu.to_device(b), cu.to_device(j), cu.alloc(8 * j.size)
        cu.pixelcmp_batch(depth, op, dA, dB, dJ, j.size, dO)
        np.testing.assert_array_equal(dO.download(np.uint64), want, err_msg=op)
        for d in (dA, dB, dJ, dO):
            d.free()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("N", [4, 8, 16, 32])
def test_intra_allangs_batch(cu, depth, N):
    O = load_oracle(depth)
    rng = np.random.default_rng(N * 3 + depth)
    dt = pixel_dtype(depth)
    n = 67
    nbp = 4 * N + 1
    nb = rng.integers(0, 1 << depth, (n, nbp)).astype(dt)
    filt = np.zeros_like(nb)
    for t in range(n):
        O.orc_intra_filter(ptr(nb, t * nbp), ptr(filt, t * nbp), N)
    dNB = cu.to_device(nb); dF = cu.alloc(nb.nbytes)
    cu.check(cu.L.x265cu_intra_filter_batch(cu.ctx, depth, N, dNB.ptr, dF.ptr, nbp, n))
    np.testing.assert_array_equal(dF.download(dt).reshape(n, nbp), filt)
    want = np.zeros((n, 33 * N * N), dt)
    for t in range(n):
        O.orc_intra_pred_allangs(ptr(want, t * 33 * N * N), ptr(nb, t * nbp), ptr(filt, t * nbp), 1, N)
    dP = cu.alloc(want.nbytes)
    cu.check(cu.L.x265cu_intra_allangs_batch(cu.ctx, depth, N, dNB.ptr, dF.ptr, nbp, dP.ptr, 1, n))
    np.testing.assert_array_equal(dP.download(dt).reshape(n, 33 * N * N), want)
    for d in (dNB, dF, dP):
        d.free()
