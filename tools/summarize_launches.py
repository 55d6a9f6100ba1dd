"""Summarises an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv [steps] [skip_launches]

`steps` divides the totals (launches captured over that many sampler steps), `skip_launches` drops the warm-up head."""
import csv, re, sys
from collections import defaultdict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
rows = rows[skip:]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("void ", "")
    v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[iu], 1e-6)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"# {len(rows)} launches, total {total / steps:.1f} ms/step over {steps:g} step(s)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v / steps:10.3f} ms/step {100 * v / total:5.1f}%  launches/step={cnt[k] / steps:6.1f}  {k}")
