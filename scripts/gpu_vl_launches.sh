#!/bin/bash
# per-launch device times of ONE vl_decode call (1 object, 1344x1344): ncu launch list of the bench script, last call only
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vl_launches_r02.csv \
   python scripts/vl_decode_bench.py --objects 1 --iters 1 > gpurun_out/vl_ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/vl_launches_r02.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
ls = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
# the last vl_decode call starts at the last rmsnorm-before-input-projection; take the launches after the last 'build_queries' minus 4
last = max(i for i, (n, _) in enumerate(ls) if "build_queries" in n)
call = ls[last - 4:]
print("launches in the last vl_decode call:", len(call), "sum of device times (us):", round(sum(t for _, t in call) / 1000, 1))
agg = {}
for n, t in call:
    k = n.split("(")[0][-60:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{t/1000:8.1f} us {c:3d}x {k}")
PY
tail -2 gpurun_out/vl_ncu.log
