#!/bin/bash
# Collects the round's measurements on one B200 into gpurun_out/r02/ (summaries are copied into profiles/ by hand).
# An ncu run only follows the same command having exited 0 without ncu.
O=gpurun_out/r02; mkdir -p $O
python bench.py --impl reference --steps 120 --warmup 30 > $O/bench_ref.json 2>$O/bench_ref.err
python bench.py --steps 120 --warmup 30 > $O/bench.json 2>$O/bench.err
python bench.py --steps 20 --warmup 5 > $O/bench_20_5.json 2>>$O/bench.err
python bench.py --config 3 --steps 60 --warmup 10 > $O/bench_cfg3.json 2>>$O/bench.err
python bench.py --config 5 --steps 60 --warmup 10 > $O/bench_cfg5.json 2>>$O/bench.err
python bench.py --stagger none --steps 120 --warmup 30 --skip-e2e --no-cpu-baseline > $O/bench_sync.json 2>>$O/bench.err
python scripts/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
[ -f build/libtmg_prof.so ] && TMG_B200_LIB=$PWD/build/libtmg_prof.so python scripts/profile_env_cycles.py --stagger none > $O/env_cycles.txt 2>&1
A="--steps 20 --warmup 10 --skip-e2e --skip-rollout --skip-no-reset --no-cpu-baseline"
python bench.py $A > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 150 --csv --log-file $O/launches.csv python bench.py $A > $O/ncu_launches.log 2>&1
A="--steps 10 --warmup 30 --skip-e2e --skip-rollout --skip-no-reset --no-cpu-baseline"
python bench.py $A > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_work|k_gate|k_pregen" -s 60 -c 6 -o $O/kernels python bench.py $A > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log; ls -la $O
