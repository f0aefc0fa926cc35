#!/bin/bash
# Batch-size sweep (BASELINE.json configs[3]): total batches of 1K..16M envs of config 2 sharded by env index over G GPUs of
# one box, one bench.py line (with clocks) per size.  usage: scripts/sweep.sh <G> [out.jsonl] [sizes...]
G=${1:-1}; OUT=${2:-gpurun_out/sweep_${G}gpu.jsonl}; shift 2 2>/dev/null
SIZES=${@:-1024 4096 16384 65536 262144 1048576 4194304 16777216}
cd "$(dirname "$0")/.."
: > "$OUT"
for total in $SIZES; do
  per=$((total / G)); [ $per -lt 1 ] && continue
  steps=60; [ $total -ge 4194304 ] && steps=20
  args="--gpus $G --steps $steps --warmup 10 --envs-per-gpu $per --skip-e2e --skip-rollout --skip-no-reset --no-cpu-baseline"
  if [ "$G" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py $args 2>/dev/null | tail -n 1 >> "$OUT"
  else
    python bench.py $args 2>/dev/null | tail -n 1 >> "$OUT"
  fi
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    c = d["config"]
    print(f"{d['n_gpus']} GPU  total envs {c['global_envs']:>9}  {d['value']/1e6:9.1f} M env-steps/s  {d['ms_per_step']:.4f} ms/step  "
          f"{d['roofline']['achieved']:.0f} GB/s per GPU ({100*d['roofline']['frac']:.2f} % of HBM)  clocks {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")
PY
