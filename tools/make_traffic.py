"""profiles/sk_traffic.json from an ncu CSV of the sk launches of one U-Net evaluation (bench.py quotes it as
`roofline.traffic` only while the kernel sources still hash to `csrc_sha`).
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sk_kernel \
       -s 123 -c 123 --csv --log-file gpurun_out/sk_traffic.csv python tools/ncu_step.py --steps 2
   python tools/make_traffic.py gpurun_out/sk_traffic.csv [out.json]"""
import csv, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import csrc_sha

src = sys.argv[1]
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "sk_traffic.json")
lines = [l for l in open(src) if not l.startswith("==")]
per = {}
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3}.get(u, 1)
    d = per.setdefault(row["ID"], {})
    d[row["Metric Name"]] = v * scale
n = len(per)
dram = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in per.values())
t = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
out = {"kernel": "sk_kernel", "launches": n, "dram_bytes_per_launch_avg": dram / max(n, 1), "dram_bytes_per_eval": dram,
       "ncu_time_per_launch_avg_us": t / max(n, 1), "csrc_sha": csrc_sha(), "batch": 16, "length": 262144, "precision": "bf16",
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sk_kernel "
              "-s 123 -c 123 (the 123 sk launches of the second U-Net evaluation, B=16, L=262144, bf16); tools/make_traffic.py"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
