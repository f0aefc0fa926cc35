#!/bin/bash
# Collects the G-GPU lines of the round on one box: bench.py (config 2 and 3) under torchrun and the batch sweep.
G=$1; O=gpurun_out/r02; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $G "$@"; }
run --steps 120 --warmup 30 > $O/bench_${G}gpu.json 2> $O/bench_${G}gpu.err
run --config 3 --steps 60 --warmup 10 > $O/bench_cfg3_${G}gpu.json 2>> $O/bench_${G}gpu.err
scripts/sweep.sh $G $O/sweep_${G}gpu.jsonl ${SWEEP_SIZES:-}
python - <<PY
import json
for f in ("bench_${G}gpu.json", "bench_cfg3_${G}gpu.json"):
    d = json.loads(open("$O/" + f).read().strip().splitlines()[-1])
    print(f, "value %.1fM (%.1fM per GPU) e2e %.1fM (%.1fM per GPU) frac %.4f" % (d["value"] / 1e6, d["value"] / 1e6 / d["n_gpus"], d["e2e"]["value"] / 1e6, d["e2e"]["value"] / 1e6 / d["n_gpus"], d["roofline"]["frac"]), d["config"]["global_envs"])
PY
