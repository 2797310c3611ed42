#!/bin/bash
# tuning helper: bench every library variant under jitterbug_b200/_lib/variants/ (built locally with -D flags)
for d in jitterbug_b200/_lib/variants/*/; do
  v=$(basename $d)
  JB_LIB=$d/libjbsim.so python bench.py --no-cpu-baseline --steps ${STEPS:-6} --burn-in ${BURN:-100} "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['config']; print('$v', round(d['value']/1e6,3), 'M steps/s', round(d['ms_per_step'],2), 'ms', 'ncon', round(c['contacts_per_substep'],3), 'newton', round(c['newton_iters_per_substep'],3))"
done
