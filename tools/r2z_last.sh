#!/bin/bash
# tools/r2z_last.sh: what the last GPU minutes of round 2 go to -- the full bench line of the final build, a small `ncu --set full`
# capture of the kernels this session changed (8 pairs per launch so that ncu's memory save / restore stays cheap), the launch list
# of the bench command.  Nothing large stays in gpurun_out/ (the 48-pair capture of tools/r2y_final.sh was 260 MB and was not copied back).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
tail -c 300 gpurun_out/r2z_bench_n1.json
OFB_BATCH=8 timeout 100 ncu --set full --clock-control none -k regex:'k_um0|k_iter64|k_polyexp2' --launch-skip 3 -c 15 -f -o /tmp/r2z_small \
    python tools/profile_pair.py 8 > gpurun_out/r2z_ncu_full.log 2>&1
timeout 40 ncu -i /tmp/r2z_small.ncu-rep --page raw --csv > gpurun_out/r2z_raw.csv 2>/dev/null
ls -la /tmp/r2z_small.ncu-rep gpurun_out/r2z_raw.csv
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2z_bench_launch_list.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-latency --no-rough > gpurun_out/r2z_bench_under_ncu.log 2>&1
ls -la gpurun_out | head -20
