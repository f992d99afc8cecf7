out=gpurun_out; tag=r02w
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x > $out/${tag}_parity.txt 2>&1; echo "parity rc=$?"; tail -4 $out/${tag}_parity.txt | cut -c1-400
run() { echo "== $1 $2 [$3]"; env $3 timeout 300 python tools/build_once.py $1 $2 7 $4 2>&1 | tail -1; }
run 27 16 "ORB_X=0"
run 27 16 "ORB_PAR_FINISH=0"
run 26 14 "ORB_X=0" gaussian
run 26 14 "ORB_X=0" plummer
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 16 2 > $out/${tag}_c3_levels.txt 2>&1
