#!/bin/bash
# One gpurun call: GPU tests, bench line, launch list of the bench command, ncu --set full of every default kernel.
# usage (on the GPU box, from the repo root): bash tools/run_round_profiles.sh <tag>
TAG=${1:-vX}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?" 
tail -3 $O/gpu_tests_$TAG.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_${TAG}_reference_arm.json 2>> $O/bench_$TAG.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-train --no-per-config --no-cpu > $O/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
N=$(python tools/prof_all.py 1 | awk '/library launches/ {print $NF}')
echo "launches per pass: $N"
ncu --set full --clock-control none --import-source on -k regex:'corr_|roipool_|psb|psroipool_|gemm_|th_' -s $N -c $N \
    -f -o $O/prof_all_$TAG python tools/prof_all.py 2 > $O/prof_all_$TAG.log 2>&1; echo "ncu full rc=$?"
python tools/time_ops.py > $O/time_ops_$TAG.txt 2>&1; echo "time_ops rc=$?"
