#!/bin/bash
# usage (on the GPU box, from the repo root): tools/profile_round.sh <tag>
# 1. plain bench (exit 0 required)  2. ncu launch list of the same command  3. ncu --set full of one launch of each hot kernel
tag=${1:-r1}
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -5 gpurun_out/bench_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-aer --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/prof_run.py 256 128 50 1 > gpurun_out/prof_plain_$tag.log 2>&1 || { tail -5 gpurun_out/prof_plain_$tag.log; exit 1; }
ncu --set full --import-source on --clock-control none --kernel-name regex:"k_sw_solve|k_lw_solve|k_sw_reduce|k_lw_reduce" -c 4 \
    -o gpurun_out/prof_${tag}_full -f python tools/prof_run.py 256 128 50 1 > gpurun_out/prof_ncu_$tag.log 2>&1
tail -2 gpurun_out/prof_ncu_$tag.log
for k in k_lw_sweep k_sw_sweep; do
  ncu --set full --clock-control none --kernel-name $k -c 3 -o gpurun_out/prof_${tag}_$k -f python tools/prof_run.py 256 128 50 1 > gpurun_out/prof_ncu_${k}_$tag.log 2>&1
  tail -1 gpurun_out/prof_ncu_${k}_$tag.log
done
