"""torch.distributed plumbing for the multi-GPU legs (one process per GPU).

The data path has no collective: ranks only agree on timing (barrier + max over ranks) and on the
totals they report.  Backend is NCCL on GPUs and gloo in the CPU tests.
"""
import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialises the default process group from RANK / WORLD_SIZE / MASTER_* when WORLD_SIZE > 1."""
    rank, local_rank, world = env_rank()
    if world <= 1:
        return rank, local_rank, world
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs closest to its GPU (NVML's ideal CPU affinity) BEFORE pinned host buffers are
    allocated, so that they are first-touched on the GPU's NUMA node.  With several ranks per box the frames and
    pictures of each rank then cross one PCIe root complex and one memory controller, not the socket interconnect.
    Returns the number of CPUs bound to, or 0 when NVML / affinity control is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def _device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_max(x):
    """max over ranks of a Python float (identity when not distributed)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_ints(values):
    """all_gather of a small list of ints -> list (per rank) of lists."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [list(values)]
    import torch
    t = torch.tensor(list(values), dtype=torch.int64, device=_device())
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
