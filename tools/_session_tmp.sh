out=gpurun_out; tag=r02u
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > $out/${tag}_gpu_tests.txt 2>&1; echo "gpu tests rc=$?"; tail -4 $out/${tag}_gpu_tests.txt | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { echo "== $1 $2 [$3]"; env $3 timeout 300 python tools/build_once.py $1 $2 9 2>&1 | tail -1; }
run 27 16 "ORB_X=0"
run 24 12 "ORB_X=0"
run 26 14 "ORB_X=0"
ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 16 2 > $out/${tag}_c3_levels.txt 2>&1
