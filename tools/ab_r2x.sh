#!/bin/bash
# tools/ab_r2x.sh: one GPU call of round 2's last session -- the full GPU test suite on the default library (packed f32x2
# UpdateMatrices, 4 rows per k_um0 thread, interior-specialised S phase), then bench.py per library variant.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests, default library"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2x_gpu_tests.log
export AB_BENCH_ARGS="${AB_BENCH_ARGS:-}"
for v in "$@"; do
  if [ "$v" = default ]; then unset OFB_LIB_PATH; else export OFB_LIB_PATH=$PWD/optical_flow_b200/lib/libofb200_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-latency --no-rough $AB_BENCH_ARGS > gpurun_out/r2x_ab_$v.log 2>&1
  python - "$v" <<'PY' | tee -a gpurun_out/r2x_ab.log
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r2x_ab_%s.log" % v).read().strip().splitlines()[-1])
    legs = d.get("legs", {})
    print(v, "proto", round(d["value"]), "dev", round(d["device_resident"]["value"]), "fast", round(legs.get("fast_arithmetic", {}).get("value", 0)),
          {k: round(x["total_ms"], 2) for k, x in d["kernels"].items() if x["total_ms"] > 0.25}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(v, "FAILED", e); print(open("gpurun_out/r2x_ab_%s.log" % v).read()[-1500:])
PY
done
for v in "$@"; do
  if [ "$v" != default ] && [ "$v" != base ]; then
    export OFB_LIB_PATH=$PWD/optical_flow_b200/lib/libofb200_$v.so
    echo "parity $v: $(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k 'stage or reference_parameters or pair_and_shot' 2>&1 | tail -1)" | tee -a gpurun_out/r2x_ab.log
  fi
done
