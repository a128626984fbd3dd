"""Per-source-line stall-sample ranking from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; si = r.index("# Samples"); ii = r.index("Instructions Executed"); continue
    if hdr and len(r) > si and r[0].isdigit() and r[si].isdigit():
        stall = {hdr[k]: int(r[k]) for k in range(len(hdr)) if hdr[k].startswith("stall_") and "Not Issued" not in hdr[k] and r[k].isdigit() and int(r[k]) > 0}
        out.append((int(r[si]), cur, int(r[0]), r[1].strip()[:110], int(r[ii]), stall))
tot = sum(o[0] for o in out)
print("total samples", tot)
for s, f, ln, src, inst, stall in sorted(out, key=lambda o: -o[0])[:topn]:
    top = sorted(stall.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*s/tot:5.1f}% {f}:{ln:<5d} inst={inst:<10d} {' '.join(f'{k[6:]}={v}' for k, v in top):40s} | {src}")
