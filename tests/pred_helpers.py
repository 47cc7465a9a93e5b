"""Prediction-cost jobs (AMVP candidate SAD, merge candidate / bidir SATD: search.cpp:1901-2023, 2474-2607): the oracle's
job struct (oracle/oracle_pred.c), the reference shim call (oracle/ref_shim.cpp: x265ref_pred_cost) and a case generator
shared by the CPU pin test and the GPU parity test."""
import ctypes as C

import numpy as np

from common import P, I, IP, pixel_dtype
from me_helpers import make_plane

LUMA_PUS = [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64), (8, 4), (4, 8),
            (16, 12), (12, 16), (16, 4), (4, 16), (32, 24), (24, 32), (32, 8), (8, 32), (64, 48), (48, 64), (64, 16), (16, 64)]


class OrcPredJob(C.Structure):
    _fields_ = [("fenc", P * 3), ("ref0", P * 3), ("ref1", P * 3), ("stride", IP), ("cstride", IP), ("w", I), ("h", I),
                ("mv0", I * 2), ("mv1", I * 2), ("cost", I), ("chroma", I), ("biAvgPP", I)]


def make_planes(depth, rng, smooth, nrefs=2, W=256, H=192, margin=96):
    """Source + `nrefs` reference pictures, 4:2:0 (chroma at half size with half the margin)."""
    mx = (1 << depth) - 1
    fenc, stride, org = make_plane(rng, depth, W, H, margin, smooth=smooth)

    def ref_of(src, dy, dx):
        if not smooth:
            return rng.integers(0, mx + 1, src.shape).astype(src.dtype)
        sh = np.roll(np.roll(src, dy, axis=0), dx, axis=1)
        return np.clip(sh.astype(np.int64) + rng.integers(-3, 4, sh.shape), 0, mx).astype(src.dtype)
    shifts = [(int(rng.integers(-5, 6)) * 2, int(rng.integers(-5, 6)) * 2) for _ in range(nrefs)]
    refs = [ref_of(fenc, *s) for s in shifts]
    fc, rc = [], [[] for _ in range(nrefs)]
    for _ in range(2):
        c, cstride, corg = make_plane(rng, depth, W // 2, H // 2, margin // 2, smooth=smooth)
        fc.append(c)
        for r in range(nrefs):
            rc[r].append(ref_of(c, shifts[r][0] // 2, shifts[r][1] // 2))
    return dict(fenc=fenc, refs=refs, fc=fc, rc=rc, stride=stride, cstride=cstride, org=org, corg=corg, W=W, H=H, margin=margin, depth=depth)


def random_job(pl, rng, w, h, kind):
    """kind: 'amvp' (SAD, one list), 'merge_uni', 'merge_bi' (SATD + chroma, addAvg), 'bidir_pp' (SATD, pixelavg_pp)."""
    W, H = pl["W"], pl["H"]
    bx = int(rng.integers(0, (W - w) // 4 + 1)) * 4; by = int(rng.integers(0, (H - h) // 4 + 1)) * 4
    bx &= ~1; by &= ~1
    mvs = [(int(rng.integers(-120, 121)), int(rng.integers(-120, 121))) for _ in range(2)]
    if rng.integers(0, 4) == 0:
        mvs[0] = (mvs[0][0] & ~3, mvs[0][1] & ~3)               # full-pel
    if rng.integers(0, 4) == 0:
        mvs[1] = (mvs[1][0] & ~3, mvs[1][1])                    # vertical only
    if rng.integers(0, 4) == 0:
        mvs[0] = (mvs[0][0], mvs[0][1] & ~3)                    # horizontal only
    lists = {"amvp": [int(rng.integers(0, 2))], "merge_uni": [int(rng.integers(0, 2))], "merge_bi": [0, 1], "bidir_pp": [0, 1]}[kind]
    return dict(bx=bx, by=by, w=w, h=h, mvs=mvs, lists=lists, cost=0 if kind == "amvp" else 1,
                chroma=int(kind in ("merge_uni", "merge_bi")), biAvgPP=int(kind == "bidir_pp"),
                ref=[int(rng.integers(0, len(pl["refs"]))), int(rng.integers(0, len(pl["refs"])))])


def plane_ptrs(pl, jb, es):
    """Pointers (ints) of source / list-0 / list-1 planes at the PU origin; an unused list is None."""
    off = pl["org"] + jb["by"] * pl["stride"] + jb["bx"]
    coff = pl["corg"] + (jb["by"] >> 1) * pl["cstride"] + (jb["bx"] >> 1)
    fenc = [pl["fenc"].ctypes.data + off * es, pl["fc"][0].ctypes.data + coff * es, pl["fc"][1].ctypes.data + coff * es]
    out = [fenc]
    for l in (0, 1):
        if l in jb["lists"]:
            r = jb["ref"][l]
            out.append([pl["refs"][r].ctypes.data + off * es, pl["rc"][r][0].ctypes.data + coff * es, pl["rc"][r][1].ctypes.data + coff * es])
        else:
            out.append(None)
    return out


def oracle_cost(O, pl, jb):
    es = np.dtype(pixel_dtype(pl["depth"])).itemsize
    fenc, r0, r1 = plane_ptrs(pl, jb, es)
    j = OrcPredJob()
    for k in range(3):
        j.fenc[k] = fenc[k]; j.ref0[k] = r0[k] if r0 else None; j.ref1[k] = r1[k] if r1 else None
    j.stride, j.cstride, j.w, j.h = pl["stride"], pl["cstride"], jb["w"], jb["h"]
    j.mv0[0], j.mv0[1] = jb["mvs"][0]; j.mv1[0], j.mv1[1] = jb["mvs"][1]
    j.cost, j.chroma, j.biAvgPP = jb["cost"], jb["chroma"], jb["biAvgPP"]
    O.orc_pred_cost.argtypes = [C.POINTER(OrcPredJob)]; O.orc_pred_cost.restype = C.c_int
    return int(O.orc_pred_cost(C.byref(j)))


def ref_cost(R, pl, jb):
    es = np.dtype(pixel_dtype(pl["depth"])).itemsize
    fenc, r0, r1 = plane_ptrs(pl, jb, es)
    arr = lambda v: (P * 3)(*v) if v else None
    mv0 = (I * 2)(*jb["mvs"][0]); mv1 = (I * 2)(*jb["mvs"][1])
    R.x265ref_pred_cost.restype = C.c_int
    R.x265ref_pred_cost.argtypes = [C.POINTER(P), C.POINTER(P), C.POINTER(P), IP, IP, I, I, C.POINTER(I), C.POINTER(I), I, I, I]
    return int(R.x265ref_pred_cost(arr(fenc), arr(r0), arr(r1), pl["stride"], pl["cstride"], jb["w"], jb["h"], mv0, mv1,
                                   jb["cost"], jb["chroma"], jb["biAvgPP"]))
