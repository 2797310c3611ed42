#!/bin/bash
# tuning helper: steady-state step time of the default library and of every variant under jitterbug_b200/_lib/variants/
# (python -m jitterbug_b200.build --variant NAME -DFLAG...).  usage: gpu_variants_time.sh [N] [burn] [K]
N=${1:-1048576}; B=${2:-150}; K=${3:-10}
echo "default: $(python tests/scripts/gpu_time.py $N $B $K)"
for d in jitterbug_b200/_lib/variants/*/; do
  v=$(basename $d)
  echo "$v: $(JB_LIB=$d/libjbsim.so python tests/scripts/gpu_time.py $N $B $K 2>&1 | tail -1)"
done
