#!/bin/bash
# Final build (default = exact arithmetic) on one 8-GPU box: the default bench at N = 8 / 4 / 2 with all legs, configs[3] and the
# sharded shot with JPEG delivery.  gpurun --gpus 8 --timeout 1200 -- tools/r2_multigpu_final.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29600
run() {
  n=$1; out=$2; shift 2
  port=$((port + 1))
  if [ "$n" = 1 ]; then timeout 600 python "$@" >> "$out" 2>> gpurun_out/r2v_err.log
  else timeout 600 $TR --nproc-per-node $n --master-port $port "$@" >> "$out" 2>> gpurun_out/r2v_err.log; fi
  echo "[$n] $* -> $(tail -c 200 "$out" | tr '\n' ' ' | cut -c1-200)"
}
for n in 8 4 2; do run $n gpurun_out/r2v_bench.jsonl bench.py --gpus $n --steps 6 --warmup 3 --no-cpu-baseline --no-latency --no-parity; done
for n in 1 8; do run $n gpurun_out/r2v_long_video_jpeg.jsonl bench.py --gpus $n --workload long_video --deliver jpeg; done
run 8 gpurun_out/r2v_long_video_raw.jsonl bench.py --gpus 8 --workload long_video
for n in 2 8; do run $n gpurun_out/r2v_sharded_shot_jpeg.jsonl bench.py --gpus $n --workload sharded_shot --deliver jpeg --steps 5 --warmup 2; done
tail -3 gpurun_out/r2v_err.log | cut -c1-300
