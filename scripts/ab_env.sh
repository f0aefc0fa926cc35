#!/bin/bash
# Diagnostics: bench.py under different environment settings / arguments, one line each.  usage: scripts/ab_env.sh "<common args>" "ENV=.. -- extra args" ...
common=$1; shift
for rep in 1 2; do
for spec in "$@"; do
  envs=${spec%%--*}; extra=${spec#*--}
  env $envs python bench.py $common $extra --skip-e2e --skip-rollout --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
nr=d.get('value_no_reset',{}).get('value',0)
print('[$spec] value %.1fM  no_reset %.1fM  step_ms med %.4f max %.4f drain %.3f'%(d['value']/1e6, nr/1e6, d['step_ms']['median'], d['step_ms']['max'], d['drain_ms']))"
done; done
