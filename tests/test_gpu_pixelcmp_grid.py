"""GPU parity: x265cu_pixelcmp_grid (whole-plane grids of equal blocks, include/x265_b200.h) against the
oracle's per-block sad / satd / sa8d / sse_pp (pixel.cpp:40-377) -- aligned and displaced plane B, the
vector fast path and the generic fallback, 8 and 10 bit.  Bit-exact."""
import ctypes as C
import zlib
import numpy as np
import pytest

from common import load_oracle, pixel_dtype

pytestmark = pytest.mark.gpu

IP = C.c_ssize_t


@pytest.fixture(scope="module")
def cu():
    import x265_b200
    return x265_b200.load()


def oracle_grid(O, depth, op, a, ao, sa, b, bo, sb, bw, bh, nbx, nby):
    fn = getattr(O, "orc_" + op)
    fn.restype = C.c_uint64 if (op == "sse_pp" and depth != 8) else C.c_uint32
    fn.argtypes = [C.c_void_p, IP, C.c_void_p, IP, C.c_int, C.c_int]
    isz = a.itemsize
    out = np.zeros(nbx * nby, np.uint64)
    for by in range(nby):
        for bx in range(nbx):
            pa = a.ctypes.data + (ao + by * bh * sa + bx * bw) * isz
            pb = b.ctypes.data + (bo + by * bh * sb + bx * bw) * isz
            out[by * nbx + bx] = fn(pa, sa, pb, sb, bw, bh)
    return out


CASES = [
    # op, bw, bh, nbx, nby, (dx, dy) of plane B
    ("sad", 8, 8, 12, 5, (0, 0)), ("sad", 8, 8, 12, 5, (5, 3)), ("sad", 16, 16, 7, 3, (-3, 2)), ("sad", 32, 32, 3, 2, (1, -1)),
    ("sad", 64, 64, 2, 2, (2, 0)), ("sad", 16, 8, 6, 4, (16, 1)), ("sad", 8, 16, 6, 2, (-7, 0)),
    ("satd", 8, 8, 12, 5, (0, 0)), ("satd", 8, 8, 10, 3, (5, 3)), ("satd", 16, 16, 7, 3, (-1, 2)), ("satd", 32, 32, 3, 2, (3, -1)),
    ("satd", 64, 64, 2, 1, (6, 0)), ("satd", 16, 4, 5, 6, (2, 2)), ("satd", 8, 4, 6, 6, (0, 1)), ("satd", 64, 16, 2, 3, (-5, 0)),
    ("sa8d", 8, 8, 12, 5, (0, 0)), ("sa8d", 8, 8, 10, 3, (7, 3)), ("sa8d", 16, 16, 7, 3, (-1, 2)), ("sa8d", 32, 32, 3, 2, (2, -1)),
    ("sa8d", 64, 64, 2, 1, (0, 0)), ("sa8d", 16, 8, 4, 4, (1, 0)), ("sa8d", 8, 16, 4, 2, (0, 3)), ("sa8d", 32, 16, 3, 3, (-2, 0)),
    ("sse_pp", 8, 8, 12, 5, (0, 0)), ("sse_pp", 16, 16, 7, 3, (-3, 2)), ("sse_pp", 64, 64, 2, 2, (5, 0)), ("sse_pp", 32, 8, 3, 5, (1, 1)),
    # generic fallback: odd block counts / unsupported widths
    ("sad", 8, 8, 11, 3, (1, 1)), ("satd", 12, 16, 5, 2, (0, 1)), ("satd", 24, 32, 3, 2, (3, 0)), ("sa8d", 48, 64, 2, 1, (1, 0)),
    ("sad", 4, 4, 9, 3, (2, 1)), ("satd", 4, 8, 8, 3, (0, 0)),
]


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind", ["rand", "extreme"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s_%dx%d_%dx%d_%+d%+d" % (c[0], c[1], c[2], c[3], c[4], c[5][0], c[5][1]))
def test_grid(cu, depth, kind, case):
    op, bw, bh, nbx, nby, (dx, dy) = case
    O = load_oracle(depth)
    dt = pixel_dtype(depth)
    maxv = (1 << depth) - 1
    rng = np.random.default_rng(zlib.crc32(repr((case, depth, kind)).encode()))
    stride = ((nbx * bw + 48 + 31) // 32) * 32
    rows = nby * bh + 16
    if kind == "rand":
        a = rng.integers(0, maxv + 1, (rows, stride)).astype(dt)
        b = rng.integers(0, maxv + 1, (rows, stride)).astype(dt)
    else:
        a = (rng.integers(0, 2, (rows, stride)) * maxv).astype(dt)
        b = (maxv - a * rng.integers(0, 2, (rows, stride))).astype(dt)
    org = 8 * stride + 16                     # 16-pixel aligned origin inside a margin
    bo = org + dy * stride + dx
    dA, dB = cu.to_device(a), cu.to_device(b)
    dO = cu.alloc(8 * nbx * nby)
    isz = a.itemsize
    cu.pixelcmp_grid(depth, op, dA.ptr + org * isz, stride, dB.ptr + bo * isz, stride, bw, bh, nbx, nby, dO)
    got = dO.download(np.uint64)
    want = oracle_grid(O, depth, op, a, org, stride, b, bo, stride, bw, bh, nbx, nby)
    np.testing.assert_array_equal(got, want)
    for d in (dA, dB, dO):
        d.free()


def test_grid_matches_job_list(cu):
    # the grid entry point and the job-list entry point are the same function of the pixels
    from x265_b200.lib import CMP_JOB
    rng = np.random.default_rng(7)
    W, H, stride = 256, 128, 320
    a = rng.integers(0, 256, (H + 16, stride)).astype(np.uint8)
    b = rng.integers(0, 256, (H + 16, stride)).astype(np.uint8)
    dA, dB = cu.to_device(a), cu.to_device(b)
    org = 8 * stride + 32
    for op in ("sad", "satd", "sa8d", "sse_pp"):
        nbx, nby = W // 16, H // 16
        xs, ys = np.meshgrid(np.arange(nbx) * 16, np.arange(nby) * 16)
        j = np.zeros(nbx * nby, CMP_JOB)
        j["a_off"] = org + ys.ravel() * stride + xs.ravel(); j["b_off"] = j["a_off"] + stride + 3
        j["a_stride"] = stride; j["b_stride"] = stride; j["w"] = 16; j["h"] = 16
        dJ = cu.to_device(j); o1 = cu.alloc(8 * j.size); o2 = cu.alloc(8 * j.size)
        cu.pixelcmp_batch(8, op, dA, dB, dJ, j.size, o1)
        cu.pixelcmp_grid(8, op, dA.ptr + org, stride, dB.ptr + org + stride + 3, stride, 16, 16, nbx, nby, o2)
        np.testing.assert_array_equal(o1.download(np.uint64), o2.download(np.uint64))
