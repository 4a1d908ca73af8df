#!/bin/bash
# tools/ab.sh <variant> [<variant> ...]: bench.py once per library variant on this GPU box ("default" = libofb200.so),
# plus the GPU parity tests for each non-default variant.  Output: one summary line per run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset OFB_LIB_PATH; else export OFB_LIB_PATH=$PWD/optical_flow_b200/lib/libofb200_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-latency --no-rough $AB_BENCH_ARGS > gpurun_out/ab_$v.log 2>&1
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.log" % v).read().strip().splitlines()[-1])
    print(v, "proto", round(d["value"]), "wall", round(d["e2e"]["value"]), "dev", round(d["device_resident"]["value"]), {k: round(x["total_ms"], 2) for k, x in d["kernels"].items()},
          d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"]["power_w_max"])
except Exception as e:
    print(v, "FAILED", e); print(open("gpurun_out/ab_%s.log" % v).read()[-1500:])
PY
done
for v in "$@"; do
  if [ "$v" != default ] && [ "$v" != base ]; then
    export OFB_LIB_PATH=$PWD/optical_flow_b200/lib/libofb200_$v.so
    echo "parity $v: $(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1)"
  fi
done
