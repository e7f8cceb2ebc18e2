"""stdin: `ncu -i rep --page source --csv`; stdout: the source lines with the most warp-stall samples (top 40)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = i
        break
if hdr is None:
    sys.exit(0)
h = rows[hdr]
si = h.index("Source")
samp = [i for i, c in enumerate(h) if c.strip() in ("# Samples", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)")]
if not samp:
    samp = [i for i, c in enumerate(h) if "Sampl" in c][:1]
ci = samp[0]
addr = h.index("Address") if "Address" in h else (h.index("#") if "#" in h else 0)
out = []
for r in rows[hdr + 1:]:
    try:
        v = float(r[ci].replace(",", ""))
    except (ValueError, IndexError):
        continue
    if v > 0:
        out.append((v, r[addr], r[si]))
tot = sum(v for v, _, _ in out) or 1.0
out.sort(reverse=True)
print("# column: %s; total samples %.0f" % (h[ci], tot))
for v, a, s_ in out[:40]:
    print("%6.2f%%  %-10s %s" % (100 * v / tot, a, s_[:150]))
