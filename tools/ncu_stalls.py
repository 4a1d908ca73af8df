"""Warp-stall breakdown per kernel launch from `ncu -i rep --page source --csv > source.csv`:
python tools/ncu_stalls.py source.csv [top]   -> per launch: share of every stall reason, and the `top` SASS lines with most samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kern, hdr, sect = None, None, []
out = []


def flush():
    if not sect:
        return
    reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {hdr[i]: 0 for i in reasons}
    allsamp = 0
    lines = []
    for r in sect:
        try:
            n = int(r[2])
        except ValueError:
            continue
        allsamp += n
        for i in reasons:
            tot[hdr[i]] += int(r[i] or 0)
        lines.append((n, r[1].strip(), {hdr[i]: int(r[i] or 0) for i in reasons if int(r[i] or 0)}))
    inst = sum(int(r[5] or 0) for r in sect if r[5].isdigit())
    print("=" * 110)
    print(kern, " samples", allsamp, " warp-instructions", inst)
    print("  " + "  ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(allsamp, 1)) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.01 * allsamp))
    for n, src, rs in sorted(lines, key=lambda t: -t[0])[:top]:
        print("  %5.1f%%  %-70s %s" % (100.0 * n / max(allsamp, 1), src[:70], " ".join("%s:%d" % (k[6:], v) for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:3])))


for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        flush()
        kern, sect = r[1], []
    elif r[0] == "Address":
        hdr = r
    elif hdr is not None and len(r) >= len(hdr) - 1:
        sect.append(r)
flush()
