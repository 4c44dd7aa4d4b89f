"""GPU box helper: the N hottest source lines of an `ncu --page source --csv` export (by warp stall samples)."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, newline="")))
hdr_i = next(i for i, r in enumerate(rows) if any("Sampling" in c or "Samples" in c for c in r))
hdr = rows[hdr_i]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
key = next((h for h in hdr if "Warp Stall Sampling (All" in h), None) or next(h for h in hdr if "Sampl" in h)
src = next((h for h in hdr if h.strip() in ("Source", "SASS", "Source (SASS)")), hdr[1])


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


tot = sum(num(r[col[key]]) for r in data) or 1.0
data.sort(key=lambda r: -num(r[col[key]]))
print(f"# {path}: top {top} lines by '{key}' (total {tot:.0f} samples); columns: share, samples, address/line, text")
for r in data[:top]:
    print(f"{100 * num(r[col[key]]) / tot:5.1f}%  {r[col[key]]:>8s}  {r[0]:>10s}  {r[col[src]][:140]}")
