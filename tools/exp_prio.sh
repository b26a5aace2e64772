#!/bin/bash
# stream-priority / BPTT-gating sweep of the C2 step (device-resident ms/step)
python - <<'PY'
import torch
print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "?")
PY
for cfg in "-1,0,0 done,pre" "-2,-1,0 done,pre" "-1,-1,0 done,pre" "-2,-1,0 pre,pre" "-2,-1,0 none,none" "-3,-2,0 done,pre" "-2,-2,-1 done,pre"; do
  set -- $cfg
  r=$(MMDA_PRIO=$1 MMDA_ORDER=$2 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'])")
  echo "prio=$1 order=$2 -> $r"
done
