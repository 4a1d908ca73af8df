"""profiles/<round>_traffic.json from an `ncu --set full ... --page raw --csv` export of tools/profile_pair.py:
measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per frame pair and per kernel family.

    python tools/ncu_traffic.py raw.csv <pairs in the capture> "<description>" > profiles/rN_traffic.json

The capture must hold whole chunks (every launch of the pairs it names).  Kernel families are bench.py's names.
"""
import csv
import json
import sys

import re

# (regex on ncu's demangled kernel name, family) -- first match wins
FAMILY = [(r"k_polyexp2<\d+, [12]>", "polyexp_scale0"), (r"k_polyexp2<", "polyexp_level"), (r"k_pyr_vf<", "pyr_v_u8"),
          (r"k_pyr_hf<", "pyr_h"), (r"k_pyr_h", "pyr_h_u8"), (r"k_pyr_v", "pyr_v"), (r"k_um0<0>", "um0_zero"),
          (r"k_um0<2>", "um0_upsample"), (r"k_um0<1>", "um0_flow"),
          (r"k_iter64<\d+, 1>", "iter_fused"), (r"k_iter64<\d+, 0>", "iter_last"), (r"k_pyr_fused", "pyr_fused"),
          (r"k_iter<\d+, 1, \d+, 0>", "iter_fused"), (r"k_iter<\d+, 0, \d+, 0>", "iter_last"),
          (r"k_iter<\d+, 1, \d+, 1>", "iter_fused_gauss"), (r"k_iter<\d+, 0, \d+, 1>", "iter_last_gauss"),
          (r"k_flow_to_bgr", "flow_to_bgr_v4"), (r"k_minmax_reset", "minmax_reset"), (r"k_minmax", "minmax_mag"),
          (r"k_bgr2gray", "bgr2gray"), (r"k_resize_u8", "resize_u8"), (r"k_build_hsv_table", "hsv_table")]


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pairs = int(sys.argv[2])
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for d in data:
        name = d[ix["Kernel Name"]]
        name = name.replace("(int)", "").replace("(bool)", "")
        fam = next((f for key, f in FAMILY if re.search(key, name)), name[:40])
        def val(m):
            return float(d[ix[m]].replace(",", "")) * unit_scale(units[ix[m]])
        e = out.setdefault(fam, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "ncu_us": 0.0})
        e["launches"] += 1
        e["dram_read"] += val("dram__bytes_read.sum")
        e["dram_write"] += val("dram__bytes_write.sum")
        e["ncu_us"] += val("gpu__time_duration.sum")
    kernels = {}
    for fam, e in out.items():
        kernels[fam] = {"launches_in_capture": e["launches"],
                        "dram_bytes_per_pair": round((e["dram_read"] + e["dram_write"]) / pairs, 1),
                        "dram_read_per_pair": round(e["dram_read"] / pairs, 1),
                        "dram_write_per_pair": round(e["dram_write"] / pairs, 1),
                        "ncu_us_per_pair": round(e["ncu_us"] / pairs, 3)}
    total = sum(k["dram_bytes_per_pair"] for k in kernels.values())
    print(json.dumps({"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "pairs_in_capture": pairs,
                      "dram_bytes_per_pair_all_kernels": round(total, 1), "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
