"""Small-frame driver for ncu captures: python profiles/run_small.py [W H stages reps chroma]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import x265_b200
from frame_helpers import gen_luma, gen_chroma, make_field

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else 576
stages = int(sys.argv[3]) if len(sys.argv) > 3 else 7
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
chroma = int(sys.argv[5]) if len(sys.argv) > 5 else 0
lib = x265_b200.load()
an = x265_b200.Analyser(lib, W, H, depth=8, numRefs=4, method=3, subme=3, merange=57, rect=1, qp=30)
for r in range(4):
    an.set_ref(r, gen_luma(W, H, 3 - r))
an.load_inputs(gen_luma(W, H, 4), make_field(W, H, 4))
if chroma:
    an.enable_chroma()
    for r in range(4):
        an.set_ref_chroma(r, gen_chroma(W, H, 3 - r, 1), gen_chroma(W, H, 3 - r, 2))
    an.load_chroma(gen_chroma(W, H, 4, 1), gen_chroma(W, H, 4, 2))
for _ in range(reps):
    an.run_resident(stages)
    lib.sync()
    print("stage ms (me_stage, resid, intra, me_kernel):", an.stage_ms(), "jobs", an.njobs, "ME phases ms (pre, int, subpel):", ["%.2f" % v for v in lib.me_phase_ms()])
