#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 1080p Farneback frame-pairs/sec on N B200 (configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path (cv2) on host cores

A "step" is one pass of the hot path over one shot: 300 consecutive 1920x1080 frame pairs with the
reference's parameters (optical_flow.py:53-59) plus the HSV picture of every pair
(visualize_optical_flow.py:48-55).  With N > 1 (torchrun, one process per GPU) every rank processes its own
300-pair shot -- pairs are independent, there is no collective on the data path ("scaling": "weak").

Legs of the default run (one JSON line on rank 0):
  value      device-resident: the 301 u8 frames already in HBM, pictures written to HBM; CUDA events.
  e2e        the same shot through ofb_shot_host: frames in pinned host memory, H2D of every frame and D2H
             of every picture inside the timed region (CUDA events from first upload to last download).
  roofline   per-kernel CUDA-event durations (option "profile") of one extra pass; the dominant kernel's
             algorithmic bytes per launch / its average duration, against MEASURED_PEAKS.json.
  cpu_baseline  cv2.calcOpticalFlowFarneback + the four picture lines on this box's host cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "1080p Farneback frame-pairs/sec"
UNIT = "pairs/s"
TRAFFIC_JSON = "r1i_traffic.json"      # measured DRAM bytes per pair and kernel (tools/ncu_traffic.py) of the current kernels
PARAMS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)

# algorithmic HBM bytes per pixel of a launch, per kernel (SURVEY.md 8d stage model)
KERNEL_BYTES_PER_PX = {
    "polyexp_scale0": 29.0,       # scale 0: 1 read frame + (4 write + 4 read I, fused away) + 20 write R
    "polyexp_level": 24.0,        # 4 read I + 20 write R            (per level pixel)
    "um0_zero": 68.0,             # UpdateMatrices: 20 + 20 + 8 read, 20 write
    "um0_upsample": 76.0,         # + the 8 B/px flow initialisation it absorbs
    "iter_fused": 96.0,           # blur+solve (28) + UpdateMatrices (68) in one launch
    "iter_last": 28.0,            # 20 read M, 8 write flow
    "minmax_mag": 8.0,            # per frame pixel
    "flow_to_bgr_v4": 11.0,       # 8 read flow + 3 write picture
}


def workload_text(W, H, pairs):
    return ("configs[1]: %dx%d synthetic shot of %d consecutive frame pairs, pyr_scale 0.5 levels 3 winsize 15 "
            "iterations 3 poly_n 5 poly_sigma 1.2 flags 0, HSV picture per pair" % (W, H, pairs))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# --------------------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 50 ms from before the warm-up; only samples that arrived inside the timed
    window [t0, t1] are reported (all samples if the window was too short to catch one)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()

        def parse(lines):
            sm, mx, reasons, power = [], [], set(), []
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons, power

        inside = [x for x in self.lines if t0 is not None and t0 <= x[0] <= t1 + 0.06]
        window = "timed region"
        if not inside:
            inside, window = self.lines, "whole run (timed region shorter than the sampling period)"
        sm, mx, reasons, power = parse(inside)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power), "window": window}


# --------------------------------------------------------------------------------------------------
# the reference on the host cores (cv2) -- cpu_baseline leg and the --impl reference arm
# --------------------------------------------------------------------------------------------------
def _cpu_worker_init():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass


def _cpu_pair_cv2(args):
    """optical_flow.py:51-59 + visualize_optical_flow.py:48-55 on one pair (one single-threaded cv2 call)."""
    prev, nxt = args
    from oracle import cv2_reference
    bgr = cv2_reference.pair_viz(prev, nxt, **PARAMS)
    return int(bgr[0, 0, 0])


def _cpu_pair_port(args):
    prev, nxt = args
    from oracle import c_oracle
    flow = c_oracle.farneback(prev, nxt, None, **PARAMS)
    return int(c_oracle.viz(flow, 0)[0, 0, 0])


def cpu_reference_rate(frames, n_tasks, cores):
    """pairs/s of the reference's CPU path over `n_tasks` pairs on `cores` processes (cv2's Farneback loops
    are serial, SURVEY.md section 6, so one single-threaded process per core is its best configuration)."""
    import multiprocessing as mp
    try:
        import cv2
        kind, fn, ver = "reference", _cpu_pair_cv2, "cv2 " + cv2.__version__
    except Exception:
        kind, fn, ver = "port", _cpu_pair_port, "oracle/farneback_oracle.c"
        from oracle import c_oracle
        c_oracle.build()
    tasks = [(frames[i % (len(frames) - 1)], frames[i % (len(frames) - 1) + 1]) for i in range(n_tasks)]
    ctx = mp.get_context("spawn")     # never fork a process that holds cv2 threads or a CUDA context
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        pool.map(fn, tasks[:cores])                       # warm-up: imports, page-in
        t0 = time.perf_counter()
        pool.map(fn, tasks, chunksize=1)
        dt = time.perf_counter() - t0
    return n_tasks / dt, kind, ver, dt


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank, world):
    if rank != 0:
        return
    import synth_frames
    W, H = args.width, args.height
    cores = os.cpu_count() or 1
    per_step = max(cores, 8) if args.ref_pairs_per_step <= 0 else args.ref_pairs_per_step
    frames = synth_frames.shot(W, H, 9, seed=0)
    rates, dts = [], []
    for s in range(args.warmup + args.steps):
        r, kind, ver, dt = cpu_reference_rate(frames, per_step, cores)
        if s >= args.warmup:
            rates.append(r); dts.append(dt)
    total_pairs = per_step * args.steps
    value = total_pairs / sum(dts)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(dts) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(W, H, args.pairs),
                       "sample": "%d pairs per step (bounded sample of the shot) on %d host processes, "
                                 "cv2.setNumThreads(1) each" % (per_step, cores)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d x %d pairs of the %dx%d shot; %s; %s" % (args.steps, per_step, W, H, ver, cpu_model())},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# this engine
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import optical_flow_b200 as ofb
    from optical_flow_b200 import dist
    import synth_frames

    W, H, P = args.width, args.height, args.pairs
    n = W * H
    numa_cpus = dist.bind_to_gpu_numa_node(local_rank) if world > 1 else 0
    eng = ofb.Farneback(local_rank)          # raises if the CUDA library / device is missing: no fallback
    if args.batch > 0:
        eng.set_option("batch", args.batch)
    if args.batch_scale0 >= 0:
        eng.set_option("batch_scale0", args.batch_scale0)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    try:
        import torch
        torch.cuda.set_device(local_rank)
        tsync = torch.cuda.synchronize
    except Exception:
        tsync = eng.synchronize

    # synthetic shot (seeded per rank), generated straight into pinned host memory
    frames = ofb.pinned_empty((P + 1, H, W), np.uint8)
    synth_frames.shot(W, H, P + 1, seed=100 + rank, out=frames)
    bgr_host = ofb.pinned_empty((P, H, W, 3), np.uint8)

    d_frames = eng.device_alloc(frames.nbytes)
    d_bgr = eng.device_alloc(P * n * 3)
    eng.h2d(d_frames, frames)

    def barrier_sync():
        eng.synchronize(); tsync(); dist.barrier()

    # ---- leg 1: device-resident ("value") ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        eng.shot_device(d_frames, P + 1, W, H, d_bgr=d_bgr, **PARAMS)
    eng.reset_kernel_stats()
    barrier_sync()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += eng.shot_device(d_frames, P + 1, W, H, d_bgr=d_bgr, **PARAMS)
    barrier_sync()
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    clocks = sampler.stop(wall0, wall1)
    launches = sum(v[0] for v in eng.kernel_stats().values())
    t_dev = dist.reduce_max(dev_ms)
    t_wall = dist.reduce_max(wall)
    total_pairs = dist.reduce_sum(P * args.steps)
    value = total_pairs / (t_dev / 1e3)

    # sanity: the pictures of the last step are real (non-constant) results
    probe = np.empty((H, W, 3), np.uint8)
    eng.d2h(probe, d_bgr + (P - 1) * n * 3)
    assert probe.max() > 0 and probe.std() > 0, "device leg produced an empty picture"

    # ---- leg 2: end to end through the host API ("e2e") ----
    for _ in range(max(1, min(args.warmup, 2))):
        eng.shot(frames, want_bgr=True, out_bgr=bgr_host, **PARAMS)
    barrier_sync()
    e2e_ms = 0.0
    for _ in range(args.steps):
        e2e_ms += eng.shot(frames, want_bgr=True, out_bgr=bgr_host, **PARAMS)["device_ms"]
    barrier_sync()
    t_e2e = dist.reduce_max(e2e_ms)
    e2e_value = total_pairs / (t_e2e / 1e3)
    assert np.array_equal(bgr_host[P - 1], probe), "host-API picture differs from the device-resident one"

    # ---- leg 3: per-kernel CUDA-event durations (one extra pass, not part of the timed regions) ----
    roof = None
    kernels = {}
    if rank == 0:
        eng.set_option("profile", 1)
        eng.reset_kernel_stats()
        eng.shot_device(d_frames, P + 1, W, H, d_bgr=d_bgr, **PARAMS)
        stats = eng.kernel_stats()
        eng.set_option("profile", 0)
        peak, peak_src = peaks()
        sched = ofb.scale_schedule(W, H, PARAMS["pyr_scale"], PARAMS["levels"])
        sum_nk = float(sum(w * h for (_, w, h, _, _) in sched))
        n0 = float(sched[-1][1] * sched[-1][2])          # scale-0 pixels
        nK = float(sched[0][1] * sched[0][2])            # coarsest-scale pixels
        px_per_launch_total = {   # pixels processed by ALL launches of the kernel for one pair (or frame)
            "polyexp_scale0": n0, "polyexp_level": sum_nk - n0,
            "um0_zero": nK, "um0_upsample": sum_nk - nK,
            "iter_fused": sum_nk * (PARAMS["iterations"] - 1), "iter_last": sum_nk,
            "minmax_mag": float(n), "flow_to_bgr_v4": float(n),
        }
        tot_ms = sum(v[1] for v in stats.values())
        for name, (cnt, ms) in sorted(stats.items(), key=lambda kv: -kv[1][1]):
            k = {"launches": cnt, "total_ms": round(ms, 3), "share": round(ms / tot_ms, 4) if tot_ms else None}
            if name in KERNEL_BYTES_PER_PX and ms > 0:
                frames_factor = (P + 1) if name.startswith("polyexp") else P
                byts = KERNEL_BYTES_PER_PX[name] * px_per_launch_total[name] * frames_factor
                k["achieved_gbs"] = round(byts / (ms * 1e-3) / 1e9, 1)
                k["alg_bytes_per_launch"] = round(byts / cnt, 1)
            kernels[name] = k
        dom = max((kv for kv in kernels.items() if "achieved_gbs" in kv[1]), key=lambda kv: kv[1]["total_ms"], default=None)
        if dom:
            # measured DRAM bytes of the same kernel from the committed `ncu --set full` capture, per launch like `achieved`
            traffic, traffic_src = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_JSON)))
                per_pair = tj["kernels"][dom[0]]["dram_bytes_per_pair"]
                traffic = round(per_pair * P / dom[1]["launches"], 1) if (W, H) == (1920, 1080) else None
                traffic_src = "profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)" % TRAFFIC_JSON
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": dom[0], "achieved": dom[1]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(dom[1]["achieved_gbs"] / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                    "alg_bytes_per_launch": dom[1]["alg_bytes_per_launch"],
                    "avg_launch_ms": round(dom[1]["total_ms"] / dom[1]["launches"], 5),
                    "share_of_step": dom[1]["share"], "peak_source": peak_src,
                    "how": "CUDA events around every launch of one extra %d-pair pass (option profile)" % P}

    # ---- leg 4: the reference's CPU path beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ntasks = max(2 * cores, 16) if W * H <= 1920 * 1080 else max(cores, 4)
        r, kind, ver, dt = cpu_reference_rate(np.array(frames[:9]), ntasks, cores)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d pairs of the same %dx%d shot in %.1f s, %d processes x cv2.setNumThreads(1); %s; %s"
                         % (ntasks, W, H, dt, cores, ver, cpu_model())}

    if rank == 0:
        peak, peak_src = peaks()
        alg = ofb.algorithmic_bytes(W, H, with_viz=True, **PARAMS)
        per_gpu = value / world
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_text(W, H, P), "pairs_per_step_per_gpu": P,
                           "l2": "inputs larger than L2: every step streams %d MB of frames and %d MB of pictures "
                                 "plus ~300 MB of per-pair intermediates through a 126 MB L2"
                                 % (frames.nbytes // 2**20, bgr_host.nbytes // 2**20),
                           "sharding": "one %d-pair shot per GPU, no collective" % P,
                           "host_affinity": ("rank bound to %d CPUs near its GPU (NVML)" % numa_cpus) if numa_cpus else "default"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes),
                        "d2h_bytes_per_step": int(bgr_host.nbytes), "ms_per_step": t_e2e / args.steps},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": roof,
                "roofline_pipeline": {"alg_bytes_per_pair": alg, "achieved": per_gpu * alg / 1e9, "peak": peak,
                                      "unit": "GB/s", "frac": per_gpu * alg / 1e9 / peak, "peak_source": peak_src},
                "kernels": kernels,
                "cpu_baseline": cpu,
                "wall_s": t_wall}
        print(json.dumps(line), flush=True)
    eng.device_free(d_frames)
    eng.device_free(d_bgr)
    dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=300)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--ref-pairs-per-step", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (experiments)")
    ap.add_argument("--batch", type=int, default=0, help="pairs per launch inside a shot (0 = engine default)")
    ap.add_argument("--batch-scale0", type=int, default=-1, help="pairs per launch at scale 0 (-1 = default, 0 = same as --batch)")
    args = ap.parse_args()

    from optical_flow_b200 import dist
    rank, local_rank, world = dist.env_rank()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        dist.init("nccl")
    run_ours(args, rank, local_rank, world)
    dist.finalize()


if __name__ == "__main__":
    main()
