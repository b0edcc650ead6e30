#!/bin/bash
# A/B variant builds: scripts/ab_libs.sh config iters lib1 lib2 ...
CFG=$1; IT=$2; shift 2
for lib in "$@"; do
  MVD_LIB=$lib python scripts/prof_passes.py $CFG $IT | python -c "
import json,sys; d=json.load(sys.stdin)
print('$lib', d['config'], 'Gvvi/s', round(d['Gvvi_s'],2), 'ms/view', round(d['sum_ms_per_view_update'],3), ' '.join(p['pass']+':'+str(p['ms_per_launch']) for p in d['passes']))"
done
