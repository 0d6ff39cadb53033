#!/bin/bash
# N=2: side-stream all-gather (default) vs all-gather ordered before the prefill; same box, steps 6
mkdir -p gpurun_out
for mode in async sync async; do
  if [ "$mode" = "sync" ]; then export PADT_SYNC_ALLGATHER=1; else unset PADT_SYNC_ALLGATHER; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 \
     --no-cpu-baseline > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/n2_$mode.json").read().strip().splitlines()[-1])
    print("$mode: value", l["value"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "e2e ms", l["e2e"]["ms_per_step"], "decode ms", l["roofline"]["ms_per_decode_step"], "parity", l.get("sharded_parity"))
except Exception as e:
    print("$mode: no line", e)
PY
done
