"""Lookahead (BASELINE configs[1], 1920x1080 8-bit): Lowres init + intra estimate + estimateFrameCost for the P and B
triples of a short window, device vs the REAL reference classes (oracle/_ref, single host thread: the reference's
CostEstimateGroup::singleCost path is serial per TLD).  python profiles/lookahead_bench.py [--frames 8]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import x265_b200
from common import load_ref, load_oracle
from frame_helpers import gen_luma
from test_gpu_lookahead import GpuLookahead
from test_lookahead_oracle_vs_ref import RefLookahead

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=8); args = ap.parse_args()
W, H, N = 1920, 1080, args.frames
frames = [gen_luma(W, H, i) for i in range(N)]
cu = x265_b200.load()
triples_p = [(b - 1, b, b) for b in range(1, N)]
triples_b = [(b - 1, b + 1, b) for b in range(1, N - 1)]
# device
t0 = time.perf_counter(); g = GpuLookahead(cu, frames, 8); cu.sync(); t_setup = time.perf_counter() - t0
cu.timer_begin()
costs_gpu = g.cost_batch(triples_p) + g.cost_batch(triples_b)      # two launches: all P triples, then all B triples
ms = cu.timer_end()
ncu = g.ncu
print("device: lowres init+intra for %d frames (incl. uploads, host numpy padding) %.1f ms; %d frame costs (P+B) %.1f ms  -> %.0f lowres CUs/s"
      % (N, 1000 * t_setup, len(costs_gpu), ms, len(costs_gpu) * ncu / (ms / 1000)))
R = load_ref(8)
if R is not None:
    t0 = time.perf_counter(); r = RefLookahead(R, frames); t_rs = time.perf_counter() - t0
    t0 = time.perf_counter(); costs_ref = [r.cost(*t) for t in triples_p + triples_b]; t_rc = time.perf_counter() - t0
    print("reference (1 thread): lowres init+intra %.1f ms; frame costs %.1f ms -> %.0f lowres CUs/s" % (1000 * t_rs, 1000 * t_rc, len(costs_ref) * ncu / t_rc))
    print("bit-exact frame costs:", costs_ref == costs_gpu)
