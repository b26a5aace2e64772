#!/bin/bash
# stream-priority / BPTT-gating sweep of the C2 step (device-resident ms/step)
#   tools/exp_prio.sh "PRIO ORDER" ...     e.g.  tools/exp_prio.sh "-2,-2,0 pre,pre" "0,0,0 none,none"
if [ $# -eq 0 ]; then
  set -- "-2,-2,0 pre,pre" "-2,-2,0 none,none" "-1,-1,0 pre,pre" "0,0,0 pre,pre" "-2,-2,-1 pre,pre" "-2,-2,-2 pre,pre" "-2,-3,0 pre,pre" "-1,-2,0 none,none"
fi
for cfg in "$@"; do
  read -r prio order <<< "$cfg"
  r=$(MMDA_PRIO=$prio MMDA_ORDER=$order timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'], d['roofline']['launch_ms_all'])")
  echo "prio=$prio order=$order -> $r"
done
