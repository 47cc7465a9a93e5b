#!/usr/bin/env python
"""bench.py -- 2160p preset-slow CTU-analysis throughput (CTUs/s) of the B200 block-primitive path.

A step = one pass of the hot path over one 3840x2160 8-bit frame per GPU (BASELINE.json configs[2]):
all PUs (2Nx2N/2NxN/Nx2N at 64..8) x 4 references through motionEstimate (STAR, merange 57, subme 3),
then per CU fused MC -> residual -> DCT -> quant -> dequant -> IDCT -> recon -> SSE, then the 35-mode
intra SA8D search (DESIGN.md "Frame analysis workload").  2040 CTUs per frame.

  value : whole-job CTUs/s with inputs resident in HBM (kernels only, CUDA events, L2 flushed between steps)
  e2e   : same metric through the public call x265_b200.Analyser.analyse() with pinned HOST buffers,
          H2D of the frame + predictor field and D2H of all results inside the timed region
  --impl reference : the same workload on the reference's own C code (oracle/_ref: MotionEstimate class +
          C primitive table compiled from /root/reference) on all host cores, bounded sample per step.

Multi-GPU (torchrun, one process per GPU): the frames of a mini-GOP share one reference set, so each rank
analyses its own frame (weak scaling) and the only exchange is an NCCL broadcast of the newest
reconstructed reference plane from its owner rank, once per step.

  --shard rows : the other partition of SURVEY 8(e) (BASELINE configs[4]): ONE frame per step, its CTU rows dealt
          to the ranks (x265cu_analyser_run_rows), every owner broadcasting its reconstructed rows afterwards
          (strong scaling; value = CTUs of the frame / max-over-ranks time).  Not the default line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, DEPTH, NREFS, QP = 3840, 2160, 8, 4, 30
METHOD, SUBME, MERANGE, RECT = 3, 3, 57, 1          # preset slow: STAR / subme 3 / merange 57 / rect (param.cpp:478-492)
CTUS_PER_FRAME = ((W + 63) // 64) * ((H + 63) // 64)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    """Usable host cores: os.cpu_count() capped by the cgroup CPU quota (the GPU box exposes 128 logical CPUs
    but grants a quota of 16)."""
    n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    return n


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
CHROMA = False      # --chroma: 4:2:0 planes resident, chroma-SATD term of subpelCompare on (both arms)


def cpu_reference(sample_rows, threads, kind_pref="reference", steps=1, warmup=0):
    """The same workload on host cores: reference-from-source driver (oracle/_ref) or the oracle port."""
    from common import load_ref, load_oracle
    from frame_helpers import Workload, cpu_analyse, lambda_for
    from me_helpers import mvcost_table
    O = load_oracle(DEPTH)
    R = load_ref(DEPTH) if kind_pref == "reference" else None
    lib, fn, kind = (R, "x265ref_analyse_frame", "reference") if R is not None else (O, "orc_analyse_frame", "port")
    hs = min(H, 64 * sample_rows)
    wl = Workload(W, hs, depth=DEPTH, numRefs=NREFS, method=METHOD, subme=SUBME, merange=MERANGE, rect=RECT, qp=QP, chroma=CHROMA)
    tab = mvcost_table(O, lambda_for(QP, DEPTH))
    ctus = ((W + 63) // 64) * ((hs + 63) // 64)
    for _ in range(warmup):
        cpu_analyse(lib, fn, wl, tab, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = cpu_analyse(lib, fn, wl, tab, threads=threads)
    dt = time.perf_counter() - t0
    return {"ctus_per_s": ctus * steps / dt, "seconds": dt, "ctus": ctus, "kind": kind, "threads": threads,
            "sample": "top %d CTU rows (3840x%d crop) of the 2160p workload, %d refs, all PUs, %d step(s)" % (sample_rows, hs, NREFS, steps),
            "checksum": int(res["me_out"][:, 0].astype(np.int64).sum())}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_cores()
    rows = 1
    # size the per-step sample so that the whole run stays within a few minutes
    probe = cpu_reference(rows, threads, steps=1)
    per_row = probe["seconds"]
    budget = 120.0 / max(1, args.steps + args.warmup)
    rows = int(max(1, min(34, budget / max(per_row, 1e-3))))
    r = cpu_reference(rows, threads, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "2160p preset-slow CTU-analysis throughput", "value": r["ctus_per_s"], "unit": "CTUs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * r["seconds"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(), "cpu_baseline": {"value": r["ctus_per_s"], "unit": "CTUs/s", "cores": threads, "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["ctus_per_s"], "unit": "CTUs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(rows=False):
    if rows:
        c = workload_config()
        c["frames_per_step"] = "1 per step, CTU rows sharded over the GPUs (contiguous row blocks)"
        c["exchange"] = "NCCL broadcast of every owner's reconstructed CTU rows (one luma plane in total) per step (N>1 only)"
        return c
    return {"workload": "3840x2160 8-bit preset slow: full ME + sub-pel interp + DCT/quant + intra primitives on device (BASELINE configs[2])",
            "ctus_per_frame": CTUS_PER_FRAME, "refs": NREFS, "search": "star", "merange": MERANGE, "subme": SUBME, "rect": RECT, "qp": QP,
            "pu_jobs_per_frame": None, "frames_per_step": "1 per GPU", "l2": "256 MiB memset between timed steps (untimed) + >300 MB per-step working set",
            "exchange": "NCCL broadcast of one reference luma plane per step (N>1 only)",
            "frames": "rank k: frame 4+k against frames 3+k..k (frame-parallel: own nearest references per frame)"}


def plane_as_tensor(torch, ptr, nbytes, device):
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(w, device=device)


def run_ours(args, rank, world, local_rank):
    import torch
    import x265_b200
    from x265_b200 import shard
    from frame_helpers import gen_luma, gen_chroma, make_field, MARGIN_X, MARGIN_Y
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = x265_b200.load(local_rank)
    an = x265_b200.Analyser(lib, W, H, depth=DEPTH, numRefs=NREFS, method=METHOD, subme=SUBME, merange=MERANGE, rect=RECT, qp=QP)
    # synthetic clip (BASELINE.md generator).  Frame shards: rank k analyses frame 4 + k against ITS four nearest previous frames
    # (3 + k .. k), as frame-parallel encoding does, so every rank has the same temporal distances and the same search problem.
    # Row shards: one frame (4) and one reference set (3..0) for all ranks.
    rows_mode = args.shard == "rows"
    k_frame = 0 if rows_mode else shard.frame_of(0, rank, world)
    for r in range(NREFS):
        an.set_ref(r, gen_luma(W, H, NREFS - 1 - r + k_frame))
    cur = gen_luma(W, H, NREFS + k_frame)
    my_rows = shard.row_blocks(an.ctu_rows, rank, world, "block") if rows_mode else [(0, an.ctu_rows)]
    field = make_field(W, H, NREFS)
    # pinned host buffers: these are what the user hands to the public call
    pin = lib.L.x265cu_host_alloc(W * H)
    h_fenc = np.frombuffer((C.c_uint8 * (W * H)).from_address(pin), np.uint8).reshape(H, W)
    h_fenc[:] = cur
    pinf = lib.L.x265cu_host_alloc(field.nbytes)
    h_field = np.frombuffer((C.c_uint8 * field.nbytes).from_address(pinf), np.int16).reshape(field.shape)
    h_field[:] = field
    h_cb = h_cr = None
    if CHROMA:
        # the reference runs preset slow (subme 3) on a 4:2:0 source with MotionEstimate::bChromaSATD on (motion.cpp:204-212)
        an.enable_chroma()
        for r in range(NREFS):
            an.set_ref_chroma(r, gen_chroma(W, H, NREFS - 1 - r + k_frame, 1), gen_chroma(W, H, NREFS - 1 - r + k_frame, 2))
        cf = NREFS + k_frame
        pc = [lib.L.x265cu_host_alloc(W * H // 4) for _ in range(2)]
        h_cb, h_cr = [np.frombuffer((C.c_uint8 * (W * H // 4)).from_address(p), np.uint8).reshape(H // 2, W // 2) for p in pc]
        h_cb[:] = gen_chroma(W, H, cf, 1); h_cr[:] = gen_chroma(W, H, cf, 2)
        an.load_chroma(h_cb, h_cr)
    flush = lib.alloc(256 << 20)
    ref0 = incoming = None
    if world > 1:
        ptr, stride = an.recon_plane_ptr(1) if rows_mode else an.ref_plane_ptr(0)
        base = ptr - (MARGIN_Y * stride + MARGIN_X)
        ref0 = plane_as_tensor(torch, base, stride * (H + 2 * MARGIN_Y), dev)
        if not rows_mode:
            # incoming-reference plane: the owner's newest reference lands here on the other ranks (making it a reference is a
            # pointer swap the bench does not do, so that every step analyses the same frames)
            inc = lib.alloc(stride * (H + 2 * MARGIN_Y))
            incoming = plane_as_tensor(torch, inc.ptr, stride * (H + 2 * MARGIN_Y), dev)

    def exchange(step):
        if world > 1 and not rows_mode:
            shard.exchange_ref(dist, ref0, step, world, recv=incoming)     # newest reconstructed reference plane from its owner
            torch.cuda.current_stream().synchronize()

    def exchange_rows():
        # rows mode: every owner publishes the CTU rows it reconstructed (CU-size-32 recon plane)
        if world > 1 and rows_mode:
            lib.sync()
            shard.exchange_rows(dist, ref0, an.ctu_rows, world, H, an.stride, MARGIN_Y)
            torch.cuda.current_stream().synchronize()

    def analyse_e2e():
        if CHROMA:
            an.load_chroma(h_cb, h_cr)                     # the frame's chroma travels with it (H2D inside the timed region)
        for r0, r1 in my_rows:
            an.analyse_rows(h_fenc, h_field, r0, r1)

    def run_resident():
        for r0, r1 in my_rows:
            an.run_rows(r0, r1, 7)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (both paths) ----
    for s in range(args.warmup):
        exchange(s)
        analyse_e2e()
        exchange_rows()
    an.load_inputs(h_fenc, h_field)
    lib.sync()
    sampler = ClockSampler(local_rank)
    launches0 = lib.launch_count()

    # ---- resident: kernels only, per-step CUDA events, L2 flushed between steps ----
    barrier()
    sampler.start()
    res_ms, stage, phases = [], np.zeros(4), np.zeros(3)
    for s in range(args.steps):
        lib.check(lib.L.x265cu_memset(lib.ctx, flush.ptr, s & 255, flush.nbytes))
        lib.sync()
        t_ex0 = time.perf_counter()
        exchange(s)
        t_ex = (time.perf_counter() - t_ex0) * 1000.0 if world > 1 else 0.0
        lib.timer_begin()
        run_resident()
        t_k = lib.timer_end()
        t_ex0 = time.perf_counter()
        exchange_rows()
        if world > 1 and rows_mode:
            t_ex += (time.perf_counter() - t_ex0) * 1000.0
        res_ms.append(t_k + t_ex)
        stage += np.array(an.stage_ms())
        phases += np.array(lib.me_phase_ms())
    barrier()
    launches = lib.launch_count() - launches0
    t_res = float(np.sum(res_ms))

    # ---- e2e: public call with host buffers, H2D + D2H inside ----
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        exchange(s)
        analyse_e2e()
        exchange_rows()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) * 1000.0
    barrier()
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(t[0]), float(t[1])

    if rank == 0:
        stage /= args.steps
        phases /= args.steps
        units = CTUS_PER_FRAME * (1 if rows_mode else world) * args.steps
        value = units / (t_res / 1000.0)
        e2e = units / (t_e2e / 1000.0)
        peak, peak_src = peaks()
        es = 1
        plane = W * H * es
        # dominant kernel: k_me<P,2,-1>, the integer-search launch of the five motion-estimation launches.
        # Algorithmic (compulsory) bytes per launch: source plane + each reference plane
        # read once + per job the 40 B job record and the 24 B phase state read and written.
        my_jobs = sum(an.row_range(r0, r1)[1] for r0, r1 in my_rows)
        my_share = sum(r1 - r0 for r0, r1 in my_rows) / float(an.ctu_rows)
        me_bytes = int(plane * (1 + NREFS) * my_share) + my_jobs * (40 + 24 + 24)
        me_ms = phases[1]
        achieved = me_bytes / (me_ms / 1000.0) / 1e9
        # DRAM bytes of that kernel from the committed ncu capture of this same command (profiles/launches_r1.md)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "me_r1_traffic.json")))
            if "3840x2160" in tr.get("config", ""):
                traffic = float(tr["traffic_bytes_per_launch"])
        except Exception:
            pass
        cfg = workload_config(rows_mode)
        cfg["pu_jobs_per_frame"] = an.njobs
        cfg["chroma_satd"] = bool(CHROMA)
        sizes = {"resid_bytes": plane * 2 + an.ncoef * 2 + 4 * plane, "intra_bytes": plane + an.ncu * 36 * 4}
        line = {
            "metric": "2160p preset-slow CTU-analysis throughput", "value": value, "unit": "CTUs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_res / args.steps, "higher_is_better": True,
            "scaling": "strong" if rows_mode else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": cfg,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "CTUs/s", "ms_per_step": t_e2e / args.steps,
                    "h2d_bytes_per_step": int(an.h2d_bytes(field)) + (W * H // 2 if CHROMA else 0), "d2h_bytes_per_step": int(sum(an.d2h_bytes_rows(r0, r1) for r0, r1 in my_rows))},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "k_me<P,2,-1> (integer search phase of the batched motionEstimate, one warp per PU x ref)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(me_bytes), "kernel_ms": float(me_ms),
                         "note": "not HBM bound: ~380 SAD candidates per job against 88 B of compulsory traffic, planes L2-resident; ncu: 94 % of l1tex throughput (unaligned candidate-row gathers), 56-59 % of issue slots busy, DRAM traffic = 1.09 x algorithmic bytes (traffic = bytes per launch from profiles/me_r1_traffic.json); see DESIGN.md section 5, profiles/launches_r1.md, profiles/me_r1_ncu.md"},
            "stages_ms": {"me_stage": float(stage[0]), "me_prechecks": float(phases[0]), "me_integer_search": float(phases[1]), "me_subpel": float(phases[2]), "residual": float(stage[1]), "intra": float(stage[2])},
            "stage_rooflines": {
                "k_cu_residual": {"achieved": sizes["resid_bytes"] / (stage[1] / 1000.0) / 1e9, "unit": "GB/s", "frac": sizes["resid_bytes"] / (stage[1] / 1000.0) / 1e9 / peak},
                "k_intra_search": {"achieved": sizes["intra_bytes"] / (stage[2] / 1000.0) / 1e9, "unit": "GB/s", "frac": sizes["intra_bytes"] / (stage[2] / 1000.0) / 1e9 / peak}},
        }
        # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample) ----
        if world == 1 and not args.no_cpu:
            threads = host_cores()
            probe = cpu_reference(1, threads)
            rows = int(max(1, min(34, 15.0 / max(probe["seconds"], 1e-3))))
            r = cpu_reference(rows, threads) if rows > 1 else probe
            line["cpu_baseline"] = {"value": r["ctus_per_s"], "unit": "CTUs/s", "cores": threads, "kind": r["kind"], "sample": r["sample"]}
        # size-independent sanity of the full-size run: every PU job produced a result inside its search window
        mv = an.me_packed[:, 1].view(np.uint32)
        line["checks"] = {"me_cost_sum": int(an.me_packed[:, 0].astype(np.int64).sum()), "numsig_sum": int(an.cu_numsig.astype(np.int64).sum()),
                          "sse_sum": int(an.cu_sse.astype(np.uint64).sum()), "intra_best_hist_nonzero": int((an.intra_cost[:, 35] > 0).sum()),
                          "mv_nonzero": int((mv != 0).sum())}
        print(json.dumps(line), flush=True)
    an.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # keep stdout clean for the ONE JSON line: library chatter (e.g. "NCCL version ...") goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--shard", default="frames", choices=["frames", "rows"],
                    help="N>1 partition: a frame per GPU (default, weak scaling) or the CTU rows of one frame per GPU (strong scaling)")
    ap.add_argument("--chroma", action="store_true",
                    help="4:2:0 chroma planes resident and the chroma-SATD term of subpelCompare on (MotionEstimate::bChromaSATD), both arms")
    args = ap.parse_args()
    global CHROMA
    CHROMA = args.chroma
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
