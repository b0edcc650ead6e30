#!/bin/bash
# usage: scripts/sweep_env.sh VAR "v1 v2 ..." config iters   -> one compact line per value
VAR=$1; VALS=$2; CFG=${3:-c3}; IT=${4:-2}
for v in $VALS; do
  env $VAR=$v python scripts/prof_passes.py $CFG $IT | python -c "
import json,sys; d=json.load(sys.stdin)
print('$VAR=$v', d['config'], 'Gvvi/s', round(d['Gvvi_s'],2), 'ms/view', round(d['sum_ms_per_view_update'],3), ' '.join(p['pass']+':'+str(p['ms_per_launch']) for p in d['passes']))"
done
