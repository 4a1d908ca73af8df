"""Copy-only probe of the host <-> device path the shot pipeline uses (no kernels): what aggregate PCIe / host-memory
bandwidth N GPUs of one box sustain when every GPU uploads frames and downloads pictures at the same time.

    python tools/pcie_probe.py                                                    # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py    # N GPUs, one process each

Every rank owns pinned host buffers of the sizes bench.py's default workload moves per chunk (24 u8 1080p frames in,
24 BGR pictures out) and issues one cudaMemcpyAsync per chunk on its own stream, exactly as ofb_shot_host does
(optical_flow_b200/csrc/engine.cu host_impl).  Three phases, each bracketed by a barrier and timed with CUDA events on the
copy streams (max over ranks): H2D only, D2H only, both directions concurrently.  Prints one JSON line on rank 0 with
per-GPU and aggregate GB/s, and the pairs/s ceiling they imply for the raw-picture protocol (2.07 MB in + 6.22 MB out
per pair) -- the number the e2e scaling curve is to be read against.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optical_flow_b200 import dist  # noqa: E402


def main():
    rank, local_rank, world = dist.env_rank()
    if world > 1:
        dist.init("nccl")
        if "--no-bind" not in sys.argv:
            dist.bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    W, H, B = 1920, 1080, 24
    reps = 40                                                      # chunks per phase: 2 GB in, 6 GB out per rank
    h_in = torch.empty(B * W * H, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(B * W * H * 3, dtype=torch.uint8).pin_memory()
    h_in.fill_(7)
    d_in = torch.empty_like(h_in, device=dev)
    d_out = torch.full_like(h_out, 3, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def phase(up, down):
        torch.cuda.synchronize(); dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t0 = time.perf_counter()
        if up:
            with torch.cuda.stream(s_up):
                ev[0].record()
                for _ in range(reps):
                    d_in.copy_(h_in, non_blocking=True)
                ev[1].record()
        if down:
            with torch.cuda.stream(s_dn):
                ev[2].record()
                for _ in range(reps):
                    h_out.copy_(d_out, non_blocking=True)
                ev[3].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dist.barrier()
        ms_up = ev[0].elapsed_time(ev[1]) if up else 0.0
        ms_dn = ev[2].elapsed_time(ev[3]) if down else 0.0
        return dist.reduce_max(ms_up), dist.reduce_max(ms_dn), dist.reduce_max(wall)

    phase(True, True)                                              # warm-up: page-in, clocks
    res = {}
    gb_in, gb_out = reps * h_in.numel() / 1e9, reps * h_out.numel() / 1e9
    for name, (u, d) in {"h2d_only": (True, False), "d2h_only": (False, True), "both": (True, True)}.items():
        ms_up, ms_dn, wall = phase(u, d)
        r = {}
        if u:
            r["h2d_gbs_per_gpu"] = gb_in / (ms_up / 1e3); r["h2d_gbs_aggregate"] = world * gb_in / (ms_up / 1e3)
        if d:
            r["d2h_gbs_per_gpu"] = gb_out / (ms_dn / 1e3); r["d2h_gbs_aggregate"] = world * gb_out / (ms_dn / 1e3)
        if u and d:
            # pairs/s ceiling of the raw-picture protocol: the slower of the two directions bounds the pipeline
            per_pair_in, per_pair_out = W * H / 1e9, 3 * W * H / 1e9
            r["pairs_per_s_ceiling_per_gpu"] = min(r["h2d_gbs_per_gpu"] / per_pair_in, r["d2h_gbs_per_gpu"] / per_pair_out)
            r["pairs_per_s_ceiling_aggregate"] = world * r["pairs_per_s_ceiling_per_gpu"]
            r["total_gbs_aggregate"] = world * (gb_in + gb_out) / wall
        res[name] = {k: round(v, 1) for k, v in r.items()}
    if rank == 0:
        try:
            numa = sorted(os.sched_getaffinity(0))
            aff = "%d cpus (%d..%d)" % (len(numa), numa[0], numa[-1])
        except Exception:
            aff = "n/a"
        print(json.dumps({"probe": "pinned host <-> device copies, one cudaMemcpyAsync per chunk, one process per GPU",
                          "n_gpus": world, "chunk_in_mb": round(h_in.numel() / 2**20, 1), "chunk_out_mb": round(h_out.numel() / 2**20, 1),
                          "chunks_per_phase": reps, "rank0_affinity": aff, "gpu": torch.cuda.get_device_name(local_rank), **res}),
              flush=True)
    dist.finalize()


if __name__ == "__main__":
    main()
