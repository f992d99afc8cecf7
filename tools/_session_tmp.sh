out=gpurun_out; tag=r02r
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x > $out/${tag}_parity.txt 2>&1; echo "parity rc=$?"; tail -4 $out/${tag}_parity.txt | cut -c1-300
run() { echo "== $1 $2 [$3]"; env $3 timeout 300 python tools/build_once.py $1 $2 9 2>&1 | tail -1; }
run 27 16 "ORB_X=0"
run 27 16 "ORB_SAMPLE_MIN_CELL=131072"
run 27 16 "ORB_SAMPLE_MIN_CELL=262144"
run 27 16 "ORB_SAMPLE_MIN_CELL=32768"
run 24 12 "ORB_X=0"
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 16 2 > $out/${tag}_c3_levels.txt 2>&1
