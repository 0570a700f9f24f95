#!/bin/bash
# usage (on the GPU box, from the repo root): tools/profile_round.sh <tag>
# 1. plain bench (exit 0 required)  2. ncu launch list of the same command  3. ncu --set full of one launch of each hot kernel
tag=${1:-r2}
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -5 gpurun_out/bench_$tag.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-e2e --no-aer --no-cpu-baseline --no-extras > gpurun_out/plain_bench_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-aer --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/prof_run.py 256 128 50 1 > gpurun_out/prof_plain_$tag.log 2>&1 || { tail -5 gpurun_out/prof_plain_$tag.log; exit 1; }
ncu --set full --import-source on --clock-control none --kernel-name regex:"k_sw_solve|k_lw_band|k_sw_reduce|k_lw_reduce|k_mcica|k_sw_prep|k_lw_prep" -c 9 \
    -o gpurun_out/prof_${tag}_full -f python tools/prof_run.py 256 128 50 1 > gpurun_out/prof_ncu_$tag.log 2>&1
tail -2 gpurun_out/prof_ncu_$tag.log
ncu --set full --clock-control none --kernel-name regex:"k_sw_sweep" -c 6 -o gpurun_out/prof_${tag}_k_sw_sweep -f python tools/prof_run.py 256 128 50 1 > gpurun_out/prof_ncu_sweep_$tag.log 2>&1
tail -1 gpurun_out/prof_ncu_sweep_$tag.log
python tools/prof_aer.py > gpurun_out/prof_aer_plain_$tag.log 2>&1 &&
ncu --set full --import-source on --clock-control none --kernel-name regex:"k_aer" -c 2 -o gpurun_out/prof_${tag}_aer -f python tools/prof_aer.py > gpurun_out/prof_ncu_aer_$tag.log 2>&1
tail -1 gpurun_out/prof_ncu_aer_$tag.log
