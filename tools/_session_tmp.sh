out=gpurun_out; tag=r02p
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 --durations=3 > $out/${tag}_parity.txt 2>&1; echo "parity rc=$?"; tail -6 $out/${tag}_parity.txt | cut -c1-300
run() { echo "== $1 $2 [$3]"; env $3 timeout 300 python tools/build_once.py $1 $2 9 2>&1 | tail -1; }
run 27 16 "ORB_X=0"
run 24 12 "ORB_X=0"
run 27 16 "ORB_SAMPLE_Z=4"
run 27 16 "ORB_SAMPLE_Z=6"
echo "== C4g"; timeout 300 python tools/build_once.py 26 14 5 gaussian 2>&1 | tail -1
echo "== C4p"; timeout 300 python tools/build_once.py 26 14 5 plummer 2>&1 | tail -1
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 16 2 > $out/${tag}_c3_levels.txt 2>&1
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 24 12 2 > $out/${tag}_c2_levels.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py --timeout 600 --durations=3 > $out/${tag}_gpu_rest.txt 2>&1; echo "rest rc=$?"; tail -6 $out/${tag}_gpu_rest.txt | cut -c1-300
