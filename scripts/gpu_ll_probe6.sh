#!/bin/bash
mkdir -p gpurun_out
PADT_MEGA_TIMELINE=32 timeout 300 python scripts/decode_only.py --impl 1 --steps 3 --no-graph --reps 1 2>&1 | grep "mega skew" | tail -148 > gpurun_out/ll_skew.log
wc -l gpurun_out/ll_skew.log
python - <<'PY'
import re
rows=[]
for l in open("gpurun_out/ll_skew.log"):
    m=re.findall(r"-?\d+", l)
    rows.append([int(x) for x in m])
rows.sort(key=lambda r:r[4])   # by gu work
print("fastest gu:", [(r[0],r[1],r[4]) for r in rows[:6]])
print("slowest gu:", [(r[0],r[1],r[4]) for r in rows[-12:]])
import statistics
for name,idx in (("qkv",2),("o",3),("gu",4),("down",5),("head",6),("poll_qkv",7),("poll_o",8),("poll_gu",9),("poll_down",10)):
    v=[r[idx] for r in rows]; print(name, "min",min(v),"med",int(statistics.median(v)),"max",max(v))
PY
