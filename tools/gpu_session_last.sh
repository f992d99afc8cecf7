#!/bin/bash
# Last 1-GPU session of round 2: ncu launch list of the bench command and the full default bench line of the final code.
# usage (under gpurun): bash tools/gpu_session_last.sh
out=gpurun_out; tag=r02y
mkdir -p $out
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --legs '' --no-cpu-baseline > $out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 85 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; tail -2 $out/${tag}_bench.err | cut -c1-200
