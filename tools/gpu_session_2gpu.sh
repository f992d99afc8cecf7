#!/bin/bash
# One multi-GPU box session (N = number of GPUs of the box, default 2): multi-process parity tests, bench lines, per-level
# event times of the sharded C3.   usage (under gpurun --gpus N): bash tools/gpu_session_2gpu.sh <tag> <N>
tag=${1:-r02}
N=${2:-2}
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 python -m pytest tests/test_gpu_multiproc.py tests/test_gpu_orbit_cli.py -x -q -m gpu --durations=6 > $out/${tag}_mp_tests.txt 2>&1
echo "mp tests rc=$?" >> $out/${tag}_mp_tests.txt
tail -12 $out/${tag}_mp_tests.txt
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err
echo "bench rc=$?"; tail -3 $out/${tag}_bench_${N}gpu.err | cut -c1-300
ORB_MR_V1=1 timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --legs '' > $out/${tag}_bench_${N}gpu_v1.json 2> $out/${tag}_bench_${N}gpu_v1.err
echo "bench v1 rc=$?"; tail -2 $out/${tag}_bench_${N}gpu_v1.err | cut -c1-300
x=$((27 - $(python -c "print(($N).bit_length()-1)")))
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 $TR tools/mr_build_once.py $x 16 3 > $out/${tag}_c3_${N}gpu_levels.txt 2>&1
tail -2 $out/${tag}_c3_${N}gpu_levels.txt | cut -c1-300
