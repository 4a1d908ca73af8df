#!/bin/bash
# tools/ab_batch2.sh "<batch>:<batch_scale0>" ...: bench.py once per (chunk size, pairs per scale-0 launch); one summary line per run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  b=${spec%%:*}; b0=${spec#*:}
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-latency --no-rough --batch $b --batch-scale0 $b0 $AB_BENCH_ARGS > gpurun_out/abb2_${b}_${b0}.log 2>&1
  python - "$b" "$b0" <<'PY' | tee -a gpurun_out/abb2.log
import json, sys
b, b0 = sys.argv[1:3]
try:
    d = json.loads(open("gpurun_out/abb2_%s_%s.log" % (b, b0)).read().strip().splitlines()[-1])
    legs = d.get("legs", {})
    print("batch", b, "scale0", b0, "proto", round(d["value"]), "dev", round(d["device_resident"]["value"]), "jpeg", round(legs.get("jpeg", {}).get("value", 0)),
          "feature", round(legs.get("feature", {}).get("value", 0)), "fast", round(legs.get("fast_arithmetic", {}).get("value", 0)),
          {k: round(x["total_ms"], 2) for k, x in d["kernels"].items() if x["total_ms"] > 0.25}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(b, b0, "FAILED", e); print(open("gpurun_out/abb2_%s_%s.log" % (b, b0)).read()[-1500:])
PY
done
