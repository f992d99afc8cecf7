out=gpurun_out; tag=r02f
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $out/${tag}_parity.txt 2>&1; echo "parity rc=$?"; tail -3 $out/${tag}_parity.txt
echo "== 27/20 self-mode"; ORB_MR_SELF=1 timeout 300 python tools/build_once.py 27 20 3 2>&1 | tail -1
ORB_MR_SELF=1 ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 20 2 > $out/${tag}_2720_self_levels.txt 2>&1; tail -1 $out/${tag}_2720_self_levels.txt
echo "== 27/20 single"; timeout 300 python tools/build_once.py 27 20 3 2>&1 | tail -1
echo "== C3 prefuse auto / off"; timeout 300 python tools/build_once.py 27 16 5 2>&1 | tail -1; ORB_PREFUSE=0 timeout 300 python tools/build_once.py 27 16 5 2>&1 | tail -1
ORB_PREFUSE=0 ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 16 2 > $out/${tag}_c3_levels_prefuse0.txt 2>&1
echo "== C2 prefuse auto / on"; timeout 300 python tools/build_once.py 24 12 9 2>&1 | tail -1; ORB_PREFUSE=1 timeout 300 python tools/build_once.py 24 12 9 2>&1 | tail -1
