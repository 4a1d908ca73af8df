#!/bin/bash
# tools/r2y_final.sh: final single-GPU evidence of round 2 (one gpurun call): the bench line with every leg, ncu --set full of one
# 48-pair chunk (after the same command exited 0 without ncu), the launch list of the bench command, the other workloads at N = 1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
set -x
timeout 300 python tools/profile_pair.py 48 > gpurun_out/r2y_profile_pair.log 2>&1 || { tail -5 gpurun_out/r2y_profile_pair.log; exit 1; }
tail -1 gpurun_out/r2y_profile_pair.log
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2y_chunk python tools/profile_pair.py 48 > gpurun_out/r2y_ncu_full.log 2>&1
timeout 300 ncu -i gpurun_out/r2y_chunk.ncu-rep --page raw --csv > gpurun_out/r2y_raw.csv 2>/dev/null
timeout 300 ncu -i gpurun_out/r2y_chunk.ncu-rep --page source --csv > gpurun_out/r2y_source.csv 2>/dev/null
ls -la gpurun_out/r2y_chunk.ncu-rep gpurun_out/r2y_raw.csv gpurun_out/r2y_source.csv
timeout 600 python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err
tail -c 600 gpurun_out/r2y_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2y_bench_launch_list.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-latency --no-rough > gpurun_out/r2y_bench_under_ncu.log 2>&1
timeout 600 python bench.py --workload long_video --deliver jpeg > gpurun_out/r2y_long_video_jpeg_n1.json 2>> gpurun_out/r2y_err.log
timeout 600 python bench.py --workload sharded_shot --deliver jpeg --steps 5 --warmup 3 > gpurun_out/r2y_sharded_shot_jpeg_n1.json 2>> gpurun_out/r2y_err.log
tail -c 300 gpurun_out/r2y_long_video_jpeg_n1.json; tail -c 300 gpurun_out/r2y_sharded_shot_jpeg_n1.json
rm -f gpurun_out/r2y_chunk.ncu-rep.tmp
