"""Developer tool: summarise an `ncu --csv` launch list (tools/ncu_list.sh):
per-kernel totals and the per-stage launches of the block kernel."""
import csv
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)


def to_float(s):
    return float(s.replace(",", ""))


launches = OrderedDict()
for r in rows:
    d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    v = to_float(r["Metric Value"])
    unit = r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time_duration"):
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1e3)
    elif "bytes_read" in r["Metric Name"]:
        d["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif "bytes_write" in r["Metric Name"]:
        d["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

tot = defaultdict(lambda: [0.0, 0, 0.0])
for d in launches.values():
    t = tot[d["name"]]
    t[0] += d.get("us", 0)
    t[1] += 1
    t[2] += d.get("rd", 0) + d.get("wr", 0)
total = sum(t[0] for t in tot.values())
print(f"# total {total / 1e3:.1f} ms over {len(launches)} launches")
print(f"{'time_us':>12} {'count':>6} {'share':>7} {'dram_MB':>10}  kernel")
for name, t in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{t[0]:12.1f} {t[1]:6d} {100 * t[0] / total:6.2f}% {t[2] / 1e6:10.1f}  {name[:70]}")

blk = [d for d in launches.values() if d["name"].startswith("k_block_warp")]
calls, cur = [], []
for d in blk:
    if d["grid"].startswith("(1,") and cur and not cur[-1]["grid"].startswith("(1,"):
        calls.append(cur)
        cur = []
    cur.append(d)
if cur:
    calls.append(cur)
for i, c in enumerate(calls):
    print(f"# block kernel, call {i}: total {sum(d.get('us', 0) for d in c) / 1e3:.2f} ms; "
          "per stage (root -> leaves) us / DRAM MB / grid")
    print("  " + " ".join(f"{d.get('us', 0):.0f}/{(d.get('rd', 0) + d.get('wr', 0)) / 1e6:.1f}/"
                          f"{d['grid'].split(',')[0][1:]}" for d in c))
