"""Top stall SASS lines of a single-kernel .ncu-rep: python tools/ncu_sass.py file.ncu-rep [top]"""
import csv, subprocess, sys, io
f = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
S = "Warp Stall Sampling (All Samples)"
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix[S]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body if r[ix[s]].isdigit()) for s in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix[S]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls if r[ix[s]].isdigit()), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix[S]]):6d} {100*int(r[ix[S]])/max(tot,1):5.1f}% ex={r[ix['Instructions Executed']]:>8s} {r[ix['Source']].strip()[:80]:80s} {st}")
