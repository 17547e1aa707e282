"""Top stall lines of an .ncu-rep source page: python tools/ncu_src.py file.ncu-rep [kernel-id] [top]"""
import csv, subprocess, sys, io
f = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else "1"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:2])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/max(tot,1):5.1f}%  {r[ix['Source']][:90]:90s} {st}")
