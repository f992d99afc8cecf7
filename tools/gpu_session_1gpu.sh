#!/bin/bash
# One 1-GPU box session: parity suite, bench line, per-level event times of C3 with and without PreLeft.
# usage (under gpurun): bash tools/gpu_session_1gpu.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --durations=6 > $out/${tag}_parity.txt 2>&1
echo "parity rc=$?" >> $out/${tag}_parity.txt
tail -12 $out/${tag}_parity.txt
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_parity.py --durations=6 > $out/${tag}_gpu_rest.txt 2>&1
echo "rest rc=$?" >> $out/${tag}_gpu_rest.txt
tail -8 $out/${tag}_gpu_rest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; tail -3 $out/${tag}_bench.err
for pl in 1 0; do
  ORB_PRELEFT=$pl ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 16 3 > $out/${tag}_c3_levels_preleft$pl.txt 2>&1
  tail -1 $out/${tag}_c3_levels_preleft$pl.txt
  ORB_PRELEFT=$pl timeout 300 python tools/build_once.py 24 12 5 2>&1 | tail -1
done
