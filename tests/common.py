"""Shared test helpers: library loading (oracle, reference-from-source, product), PU tables, fixtures.

The reference library (oracle/_ref/libx265ref{8,10}.so) is compiled from /root/reference by
oracle/Makefile.ref; it is test infrastructure and travels to the GPU box as a prebuilt .so.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# LumaPU order of /root/reference/source/common/primitives.h:41-55
LUMA_PU = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32),
           (64, 32), (32, 64), (16, 12), (12, 16), (16, 4), (4, 16), (32, 24), (24, 32), (32, 8), (8, 32),
           (64, 48), (48, 64), (64, 16), (16, 64)]
LUMA_CU = [4, 8, 16, 32, 64]
CSP_I420 = 1

P = C.c_void_p
I = C.c_int
IP = C.c_ssize_t  # intptr_t


def pixel_dtype(depth):
    return np.uint8 if depth == 8 else np.uint16


def ref_path(depth):
    return os.path.join(ROOT, "oracle", "_ref", "libx265ref%d.so" % depth)


def oracle_path(depth):
    return os.path.join(ROOT, "oracle", "liboracle%d.so" % depth)


_cache = {}


def load_ref(depth):
    key = ("ref", depth)
    if key not in _cache:
        p = ref_path(depth)
        if not os.path.exists(p):
            return None
        L = C.CDLL(p)
        L.x265ref_get.restype = C.c_void_p
        L.x265ref_get.argtypes = [C.c_char_p, I, I, I]
        L.x265ref_dct_matrix.restype = C.c_void_p
        L.x265ref_luma_filter.restype = C.c_void_p
        L.x265ref_chroma_filter.restype = C.c_void_p
        L.x265ref_intra_filter_flags.restype = C.c_void_p
        L.x265ref_lambda.restype = C.c_double
        L.x265ref_lambda2.restype = C.c_double
        L.x265ref_lambda.argtypes = [I]
        _cache[key] = L
    return _cache[key]


def load_oracle(depth):
    key = ("orc", depth)
    if key not in _cache:
        p = oracle_path(depth)
        if not os.path.exists(p):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = C.CDLL(p)
        L.orc_dct_matrix.restype = C.c_void_p
        L.orc_var.restype = C.c_uint64
        sse = C.c_uint32 if depth == 8 else C.c_uint64
        L.orc_sse_pp.restype = sse
        L.orc_sse_ss.restype = sse
        L.orc_ssd_s.restype = sse
        L.orc_quant.restype = C.c_uint32
        L.orc_nquant.restype = C.c_uint32
        L.orc_copy_cnt.restype = C.c_uint32
        _cache[key] = L
    return _cache[key]


def ref_fn(L, name, restype, argtypes, i=0, j=0, k=0):
    """Fetch a function pointer from the reference's C table and type it."""
    ptr = L.x265ref_get(name.encode(), i, j, k)
    if not ptr:
        return None
    return C.CFUNCTYPE(restype, *argtypes)(ptr)


def ptr(a, off=0):
    """void* to element `off` of a numpy array."""
    return C.c_void_p(a.ctypes.data + off * a.itemsize)


def fixtures(rng, depth, shape, kind):
    """Harness-style inputs (pixelharness.cpp:31-80): random / all-min / all-max pixels."""
    mx = (1 << depth) - 1
    dt = pixel_dtype(depth)
    if kind == "rand":
        return rng.integers(0, mx + 1, shape).astype(dt)
    if kind == "min":
        return np.zeros(shape, dt)
    if kind == "max":
        return np.full(shape, mx, dt)
    raise ValueError(kind)


def resid_fixture(rng, depth, shape, kind):
    """Residual inputs as in mbdstharness.cpp:61-82."""
    mx = (1 << depth) - 1
    if kind == "rand":
        return (rng.integers(0, mx + 1, shape) - rng.integers(0, mx + 1, shape)).astype(np.int16)
    if kind == "min":
        return np.full(shape, -mx, np.int16)
    if kind == "max":
        return np.full(shape, mx, np.int16)
    raise ValueError(kind)


def make_plane(rng, depth, w, h, margin, kind="rand", smooth=False):
    """A picture plane with replicated margins; returns (buffer, stride, origin_offset)."""
    dt = pixel_dtype(depth)
    mx = (1 << depth) - 1
    stride = (w + 2 * margin + 63) // 64 * 64
    buf = np.zeros((h + 2 * margin, stride), dt)
    if smooth:
        yy, xx = np.mgrid[0:h, 0:w]
        img = 128 + 60 * np.sin(xx / 17.0) + 40 * np.cos(yy / 11.0) + rng.integers(-6, 7, (h, w))
        img = np.clip(img, 0, 255).astype(np.int64) << (depth - 8)
    elif kind == "rand":
        img = rng.integers(0, mx + 1, (h, w))
    else:
        img = np.full((h, w), 0 if kind == "min" else mx)
    buf[margin:margin + h, margin:margin + w] = img
    buf[margin:margin + h, :margin] = buf[margin:margin + h, margin:margin + 1]
    buf[margin:margin + h, margin + w:margin + w + margin] = buf[margin:margin + h, margin + w - 1:margin + w]
    buf[:margin, :] = buf[margin:margin + 1, :]
    buf[margin + h:, :] = buf[margin + h - 1:margin + h, :]
    return buf, stride, margin * stride + margin
