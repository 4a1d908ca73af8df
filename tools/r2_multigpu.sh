#!/bin/bash
# One 8-GPU box: copy-only probe, the default bench, configs[3] (long video) and the sharded single shot at N = 1/2/4/8.
# Every line lands in gpurun_out/r2m_*.json(l); run as:  gpurun --gpus 8 --timeout 1500 -- tools/r2_multigpu.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
run() {   # run <n> <outfile> <script and args...>
  n=$1; out=$2; shift 2
  port=$((port + 1))
  if [ "$n" = 1 ]; then timeout 600 python "$@" >> "$out" 2>> gpurun_out/r2m_err.log
  else timeout 600 $TR --nproc-per-node $n --master-port $port "$@" >> "$out" 2>> gpurun_out/r2m_err.log; fi
  echo "[$n] $* -> $(tail -c 300 "$out" | tr '\n' ' ' | cut -c1-300)"
}
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
nproc > gpurun_out/r2m_nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2m_nproc.txt; numactl -H >> gpurun_out/r2m_nproc.txt 2>&1
for n in 1 2 4 8; do run $n gpurun_out/r2m_pcie.jsonl tools/pcie_probe.py; done
for n in 8 4 2 1; do run $n gpurun_out/r2m_bench.jsonl bench.py --gpus $n --steps 6 --warmup 3 --no-cpu-baseline --no-latency --no-parity; done
for n in 1 2 4 8; do run $n gpurun_out/r2m_long_video.jsonl bench.py --gpus $n --workload long_video; done
for n in 1 2 4 8; do run $n gpurun_out/r2m_sharded_shot.jsonl bench.py --gpus $n --workload sharded_shot --steps 5 --warmup 2; done
for n in 1 8; do run $n gpurun_out/r2m_feature.jsonl bench.py --gpus $n --workload feature --steps 8 --warmup 3; done
tail -5 gpurun_out/r2m_err.log
