#!/bin/bash
# tools/ab_batch.sh <batch> ...: bench.py once per chunk size (pairs per launch inside a shot) on this GPU box; one summary line per run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-latency --no-rough --batch $b $AB_BENCH_ARGS > gpurun_out/abb_$b.log 2>&1
  python - "$b" <<'PY' | tee -a gpurun_out/abb.log
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/abb_%s.log" % v).read().strip().splitlines()[-1])
    legs = d.get("legs", {})
    print("batch", v, "proto", round(d["value"]), "wall", round(d["e2e"]["value"]), "dev", round(d["device_resident"]["value"]), "jpeg", round(legs.get("jpeg", {}).get("value", 0)),
          "feature", round(legs.get("feature", {}).get("value", 0)), "fast", round(legs.get("fast_arithmetic", {}).get("value", 0)),
          {k: round(x["total_ms"], 2) for k, x in d["kernels"].items() if x["total_ms"] > 0.25}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(v, "FAILED", e); print(open("gpurun_out/abb_%s.log" % v).read()[-1500:])
PY
done
