#!/usr/bin/env python
"""Cost of the chroma-SATD term (x265cu_me_batch_chroma vs x265cu_me_batch) on one job list: 20 000 STAR / subme 3
jobs of the preset-slow PU sizes on a 1024x576 4:2:0 picture.  CUDA-event times of whole batches; prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import x265_b200
from x265_b200.lib import ME_JOB
from common import make_plane

cu = x265_b200.load()
rng = np.random.default_rng(5)
W, H, margin, depth, n = 1024, 576, 96, 8, 20000
fenc, stride, org = make_plane(rng, depth, W, H, margin, smooth=True)
ref = np.clip(np.roll(np.roll(fenc, 3, 0), -5, 1).astype(np.int64) + rng.integers(-3, 4, fenc.shape), 0, 255).astype(np.uint8)
ch = [make_plane(rng, depth, W // 2, H // 2, margin // 2, smooth=True) for _ in range(2)]
cstride = ch[0][1]
rc = [np.clip(np.roll(np.roll(c[0], 1, 0), -2, 1).astype(np.int64) + rng.integers(-3, 4, c[0].shape), 0, 255).astype(np.uint8) for c in ch]
sizes = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8)]
jobs = np.zeros(n, ME_JOB)
for t in range(n):
    w, h = sizes[t % len(sizes)]
    bx = int(rng.integers(0, (W - w) // 8 + 1)) * 8; by = int(rng.integers(0, (H - h) // 8 + 1)) * 8
    q = (int(rng.integers(-24, 25)), int(rng.integers(-24, 25)))
    j = jobs[t]
    j["offset"] = org + by * stride + bx; j["ref"] = 0; j["pw"] = w; j["ph"] = h
    j["mvmin"] = (max(-bx - 84, (q[0] >> 2) - 57), max(-by - 84, (q[1] >> 2) - 57))
    j["mvmax"] = (min(W - w - bx + 84, (q[0] >> 2) + 57), min(H - h - by + 84, (q[1] >> 2) + 57))
    j["qmvp"] = q; j["numCand"] = 2; j["mvc"] = rng.integers(-40, 41, 8); j["method"] = 3; j["subme"] = 3; j["merange"] = 57
tab = cu.mvcost_table(11.3137, 65536)
d = dict(fenc=cu.to_device(fenc), ref=cu.to_device(ref), fcb=cu.to_device(ch[0][0]), fcr=cu.to_device(ch[1][0]),
         rcb=cu.to_device(rc[0]), rcr=cu.to_device(rc[1]), tab=cu.to_device(tab), jobs=cu.to_device(jobs), out=cu.alloc(n * 16))
t_ref = cu.to_device(np.array([d["ref"].ptr], np.uint64)); t_cb = cu.to_device(np.array([d["rcb"].ptr], np.uint64)); t_cr = cu.to_device(np.array([d["rcr"].ptr], np.uint64))
res = {}
for name in ("luma", "chroma"):
    ms, ph = [], np.zeros(3)
    for it in range(6):
        cu.timer_begin()
        if name == "luma":
            cu.me_batch(depth, d["fenc"], stride, t_ref, stride, 0, d["tab"], 65536, d["jobs"], n, d["out"])
        else:
            cu.me_batch_chroma(depth, d["fenc"], stride, t_ref, stride, d["fcb"], d["fcr"], t_cb, t_cr, cstride, d["tab"], 65536, d["jobs"], n, d["out"])
        t = cu.timer_end()
        if it >= 2:
            ms.append(t); ph += np.array(cu.me_phase_ms())
    o = d["out"].download(np.int32).reshape(n, 4)
    res[name] = {"ms": float(np.mean(ms)), "phases_ms": (ph / len(ms)).tolist(), "cost_sum": int(o[:, 0].astype(np.int64).sum())}
print(json.dumps({"jobs": n, "config": "1024x576 8-bit 4:2:0, STAR merange 57 subme 3, preset-slow PU sizes", **res}))
