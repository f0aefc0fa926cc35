#!/bin/bash
# Diagnostics: runs bench.py once per library variant in build/ (scripts/build_variant.sh) and prints value / no-reset value.
# usage: scripts/ab.sh "<bench args>" name1 name2 ...
args=$1; shift
for rep in 1 2; do
for v in "$@"; do
  TMG_B200_LIB=$PWD/build/libtmg_$v.so python bench.py $args --skip-e2e --skip-rollout --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
nr=d.get('value_no_reset',{}).get('value',0)
print('$v: value %.1fM  no_reset %.1fM  step_ms med %.4f max %.4f drain %.3f'%(d['value']/1e6, nr/1e6, d['step_ms']['median'], d['step_ms']['max'], d['drain_ms']))"
done; done
