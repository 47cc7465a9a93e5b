"""Where the c2 end-to-end step spends its host time: per-frame H2D + border extension + Lowres init, timed separately."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import x265_b200
from x265_b200.lookahead import Lookahead, MARGIN_X, MARGIN_Y
from frame_helpers import gen_luma

lib = x265_b200.load()
W, H, n = 1920, 1080, 8
frames = [gen_luma(W, H, i) for i in range(n)]
la = Lookahead(lib, W, H, 8, n)
cu = lib
for rep in range(3):
    t = {"copy2d": 0.0, "extend": 0.0, "lowres": 0.0, "sync": 0.0}
    for i, img in enumerate(frames):
        img = np.ascontiguousarray(img)
        fs = (W + 63) // 64 * 64 + 2 * MARGIN_X
        rows = (H + 63) // 64 * 64 + 2 * MARGIN_Y
        forg = (MARGIN_Y * fs + MARGIN_X) * la.es
        if la._full is None:
            la._full = cu.alloc(fs * rows * la.es)
        t0 = time.perf_counter()
        cu.check(cu.L.x265cu_copy2d(cu.ctx, la._full.ptr + forg, fs * la.es, img.ctypes.data, W * la.es, W * la.es, H, 0)); cu.sync()
        t1 = time.perf_counter()
        cu.check(cu.L.x265cu_extend_border(cu.ctx, 8, la._full.ptr + forg, fs, W, H, MARGIN_X, MARGIN_Y)); cu.sync()
        t2 = time.perf_counter()
        f = la.fr[i]
        cu.check(cu.L.x265cu_frame_init_lowres(cu.ctx, 8, la._full.ptr + forg, fs, *[p.ptr + la.lorg for p in f["planes"]], la.ls, la.w8 * 8, la.h8 * 8, MARGIN_X, MARGIN_Y)); cu.sync()
        t3 = time.perf_counter()
        t["copy2d"] += t1 - t0; t["extend"] += t2 - t1; t["lowres"] += t3 - t2
    print(rep, {k: round(1000 * v / n, 3) for k, v in t.items()}, "ms per frame", flush=True)
t0 = time.perf_counter()
for i, img in enumerate(frames):
    la.init_frame(i, img)
print("init_frame", round(1000 * (time.perf_counter() - t0) / n, 3), "ms per frame")
t0 = time.perf_counter(); la.intra_batch(list(range(n))); lib.sync(); print("intra_batch", round(1000 * (time.perf_counter() - t0), 3), "ms for", n)
