"""bench.py's c2 end-to-end step, section by section (host timers, a sync after each section)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import x265_b200, bench
from x265_b200.lookahead import Lookahead, window_triples, conflict_free_batches

bench.CFG = bench.CONFIGS["c2"]; bench.CFG_NAME = "c2"
c = bench.CFG
lib = x265_b200.load()
frames = bench.lookahead_frames()
n = len(frames)
pinned = []
for f in frames:
    p = lib.L.x265cu_host_alloc(f.nbytes)
    a = np.frombuffer((C.c_uint8 * f.nbytes).from_address(p), f.dtype).reshape(f.shape); a[:] = f; pinned.append(a)
la = Lookahead(lib, c["W"], c["H"], c["depth"], n, lookahead_slices=c.get("lslices", 0))
triples = window_triples(n, c["bframes"])
batches = conflict_free_batches(triples)
for rep in range(4):
    T = {}
    def sec(name, t0):
        lib.sync(); T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1000.0
    t0 = time.perf_counter(); la.forget_results()
    for f in la.fr: f["has_intra"] = False
    sec("forget", t0)
    t0 = time.perf_counter()
    for i in range(n): la.init_frame(i, pinned[i], sync=False)
    sec("init_frames", t0)
    t0 = time.perf_counter(); la.intra_batch(list(range(n))); sec("intra", t0)
    t0 = time.perf_counter(); preps = [la.prepare_batch(b) for b in batches]; sec("prepare", t0)
    t0 = time.perf_counter()
    for p in preps: la.launch_batch(p)
    sec("launch", t0)
    t0 = time.perf_counter()
    for p in preps: la.collect_batch(p, full=True)
    sec("collect", t0)
    print(rep, {k: round(v, 2) for k, v in T.items()}, "total", round(sum(T.values()), 1), "ms", flush=True)
