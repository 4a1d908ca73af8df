"""BASELINE.json configs[3]: a long 1080p video of many begin/end shots, sharded across the GPUs of one box.

    python tools/long_video.py [--pairs 20000] [--width 1920 --height 1080]                 # 1 GPU
    python -m torch.distributed.run --nproc-per-node N tools/long_video.py --pairs 20000      # N GPUs

Shots have random lengths of 50-400 pairs (seeded).  To bound host memory every shot draws its frames cyclically
from a bank of 65 distinct warped frames.  Whole shots are assigned to ranks greedily, longest
first (optical_flow_b200.shard_shots); a rank runs each of its shots through the HOST API (pinned frames in,
pinned pictures out, copies inside the timed region).  No collective touches the data; ranks only agree on timing.
Prints one JSON line on rank 0: aggregate pairs/s = all pairs / max-over-ranks time.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optical_flow_b200 as ofb  # noqa: E402
from optical_flow_b200 import dist  # noqa: E402
import synth_frames  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--seed", type=int, default=7)
    args = ap.parse_args()
    rank, local_rank, world = dist.env_rank()
    if world > 1:
        dist.init("nccl")
        dist.bind_to_gpu_numa_node(local_rank)
    W, H = args.width, args.height
    rng = np.random.default_rng(args.seed)
    lengths = []
    while sum(lengths) < args.pairs:
        lengths.append(int(min(rng.integers(50, 401), args.pairs - sum(lengths))))
    mine = ofb.shard_shots(lengths, world)[rank]

    eng = ofb.Farneback(local_rank)
    bank_n = 65
    bank = synth_frames.shot(W, H, bank_n, seed=1000 + rank)
    max_len = max((n for _, _, n in mine), default=1)
    frames = ofb.pinned_empty((max_len + 1, H, W), np.uint8)
    out = ofb.pinned_empty((max_len, H, W, 3), np.uint8)
    # warm-up (workspace allocation, clocks)
    frames[:17] = bank[:17]
    eng.shot(frames[:17], want_bgr=True, out_bgr=out[:16])
    eng.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    dev_ms, done, checksum = 0.0, 0, 0
    for (shot, first, n) in mine:
        # the shot's frames: a forward-backward walk through the bank, so consecutive frames always differ by one warp step
        idx = (np.arange(first, first + n + 1) + shot * 7) % (2 * bank_n - 2)
        idx = np.where(idx < bank_n, idx, 2 * bank_n - 2 - idx)
        frames[:n + 1] = bank[idx]                                   # host-side "decode" of the shot
        r = eng.shot(frames[:n + 1], want_bgr=True, out_bgr=out[:n])
        dev_ms += r["device_ms"]
        done += n
        checksum += int(out[n - 1, H // 2, W // 2].sum())
    eng.synchronize()
    wall = time.perf_counter() - t0
    dist.barrier()
    t_dev = dist.reduce_max(dev_ms)
    t_wall = dist.reduce_max(wall)
    total = dist.reduce_sum(done)
    if rank == 0:
        print(json.dumps({"workload": "configs[3]: %d pairs of %dx%d in %d shots of 50-400 pairs, sharded by whole shots"
                                      % (int(total), W, H, len(lengths)),
                          "n_gpus": world, "pairs": int(total),
                          "pairs_per_s_device_events": total / (t_dev / 1e3),
                          "pairs_per_s_wall_incl_host_frame_assembly": total / t_wall,
                          "shots_on_rank0": len(mine), "checksum_rank0": checksum}), flush=True)
    dist.finalize()


if __name__ == "__main__":
    main()
