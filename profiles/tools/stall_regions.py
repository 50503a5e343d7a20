"""Aggregates the per-instruction stall samples of an ncu SASS source page (ncu -i X.ncu-rep --page source --csv
--print-source sass) into program regions delimited by marker instructions (BAR / SYNCS / WARPSYNC / long runs), and
prints the hottest instructions.  Usage: python profiles/tools/stall_regions.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("by reason:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
# regions: split at barrier-like instructions
reg, cur, start = [], 0, 0
marks = ("BAR.", "SYNCS.", "WARPSYNC", "UTMALDG", "UBLKCP", "LDTM", "STTM", "BRA")
cum = 0
print("\n-- cumulative samples at markers --")
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]] or 0)
    cum += s
    src = r[ix["Source"]]
    if any(m in src for m in marks):
        print(f"{n:5d} {100*cum/tot:6.1f}%  +{s:6d}  ex={r[ix['Instructions Executed']]:>10}  {src[:90]}")
print("\n-- hottest instructions --")
for n, r in sorted(enumerate(data), key=lambda t: -int(t[1][ix["# Samples"]] or 0))[:top]:
    s = int(r[ix["# Samples"]] or 0)
    why = max(stalls, key=lambda k: int(r[ix[k]] or 0))
    print(f"{n:5d} {100*s/tot:5.2f}% {why:22s} {r[ix['Source']][:100]}")
