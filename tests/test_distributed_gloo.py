"""world_size-2 gloo test of the N>1 host path (no GPU): rank discovery, the barrier / max-over-ranks
timing reduction bench.py uses, and that the ranks' pair ranges tile the shot with no collective on the data."""
import os
import socket
import subprocess
import sys

from conftest import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["OFB_ROOT"])
from optical_flow_b200 import dist, shard_pairs
rank, local_rank, world = dist.init(backend="gloo")
assert world == 2
s, e = shard_pairs(301, world, rank)
dist.barrier()
t = dist.reduce_max(10.0 + rank)          # slowest rank defines the step time
total = dist.reduce_sum(e - s)
ranges = dist.gather_ints([s, e])
assert t == 11.0, t
assert total == 301, total
assert ranges == [[0, 151], [151, 301]], ranges
dist.barrier()
dist.finalize()
print("RANK_OK", rank)
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_sharding_and_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OFB_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = []
    for p in procs:
        out, _ = p.communicate(timeout=240)
        outs.append(out.decode())
        assert p.returncode == 0, out.decode()
    assert "RANK_OK 0" in outs[0] and "RANK_OK 1" in outs[1]
