#!/bin/bash
# tools/ab_opts.sh "<name>:<opt=val,opt=val>" ...: bench.py once per engine-option set on this GPU box (same library),
# one summary line per run.  Example: tools/ab_opts.sh default: f32sums:f32_window_sums=1 slowpe:polyexp_fast=0
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; opts=${spec#*:}
  args=""
  IFS=',' read -ra kv <<< "$opts"
  for o in "${kv[@]}"; do [ -n "$o" ] && args="$args --opt $o"; done
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-latency --no-rough $args $AB_BENCH_ARGS \
      > gpurun_out/abo_$name.json 2> gpurun_out/abo_$name.err
  python - "$name" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/abo_%s.json" % v).read().strip().splitlines()[-1])
    legs = d.get("legs", {})
    print(v, "proto", round(d["value"]), "dev", round(d["device_resident"]["value"]), "feature", round(legs.get("feature", {}).get("value", 0)),
          "jpeg", round(legs.get("jpeg", {}).get("value", 0)), legs.get("jpeg", {}).get("encoder_us_per_picture"), legs.get("jpeg", {}).get("encoder_kernels_ms"), {k: round(x["total_ms"], 2) for k, x in d["kernels"].items() if x["total_ms"] > 0.25},
          d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(v, "FAILED", e); print(open("gpurun_out/abo_%s.err" % v).read()[-1500:])
PY
done
