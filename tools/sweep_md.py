"""Markdown table of a tools/sweep.py result:  python tools/sweep_md.py profiles/rN_sweep.jsonl "<title>" > profiles/rN_sweep.md"""
import json
import sys

recs = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
title = sys.argv[2] if len(sys.argv) > 2 else "Parameter sweep on one B200 (tools/sweep.py)"
cores = recs[0]["cv2_cores"] if recs else 0
print("# %s, cv2 on the box's %d host cores beside it\n" % (title, cores))
print("Device-resident short shots (CUDA events), pictures included; `frac` = (B_pair+B_viz) x pairs/s / measured HBM copy "
      "bandwidth (MEASURED_PEAKS.json);\nEPE = endpoint difference vs cv2 on one pair (tolerance: mean <= 1e-3, max <= 1e-2 px).\n")
print("| size | winsize | iterations | levels | poly_n | flags | GPU pairs/s | roofline frac | EPE mean | EPE max | cv2 pairs/s (%d procs) | speed-up |" % cores)
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in recs:
    p = r["params"]
    print("| %s | %d | %d | %d | %d | %d | %.1f | %.3f | %.1e | %.1e | %.2f | %.0fx |"
          % (r["size"], p["winsize"], p["iterations"], p["levels"], p["poly_n"], p["flags"], r["gpu_pairs_per_s"], r["roofline_frac"],
             r["epe_vs_cv2_mean"], r["epe_vs_cv2_max"], r["cv2_pairs_per_s"], r["speedup"]))
