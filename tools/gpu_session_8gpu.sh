#!/bin/bash
# One 8-GPU box session: the default bench line (C3 sharded + C5 leg), the weak-scaled C2 of round 1, per-level event
# times of C5.   usage (under gpurun --gpus 8): bash tools/gpu_session_8gpu.sh <tag>
tag=${1:-r02}
N=8
out=gpurun_out
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_8gpu.json 2> $out/${tag}_bench_8gpu.err
echo "bench rc=$?"; tail -3 $out/${tag}_bench_8gpu.err | cut -c1-300
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --config c2w --legs '' > $out/${tag}_bench_8gpu_c2w.json 2> $out/${tag}_bench_8gpu_c2w.err
echo "bench c2w rc=$?"
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 $TR tools/mr_build_once.py 27 20 2 > $out/${tag}_c5_8gpu_levels.txt 2>&1
tail -2 $out/${tag}_c5_8gpu_levels.txt | cut -c1-400
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 $TR tools/mr_build_once.py 24 16 3 > $out/${tag}_c3_8gpu_levels.txt 2>&1
tail -2 $out/${tag}_c3_8gpu_levels.txt | cut -c1-400
