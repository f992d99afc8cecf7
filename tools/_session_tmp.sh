out=gpurun_out; tag=r02h
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $out/${tag}_parity.txt 2>&1; echo "parity rc=$?"; tail -3 $out/${tag}_parity.txt
for b in 1 0 1 0; do echo "== C3 bulk=$b"; ORB_PART_BULK=$b timeout 300 python tools/build_once.py 27 16 7 2>&1 | tail -1; done
for b in 1 0; do echo "== C3 bulk=$b prefuse0"; ORB_PREFUSE=0 ORB_PART_BULK=$b timeout 300 python tools/build_once.py 27 16 7 2>&1 | tail -1; done
for b in 1 0; do echo "== C2 bulk=$b"; ORB_PART_BULK=$b timeout 300 python tools/build_once.py 24 12 9 2>&1 | tail -1; done
for b in 1 0; do ORB_PART_BULK=$b ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 16 2 > $out/${tag}_c3_levels_bulk$b.txt 2>&1; done
echo "== 27/20 self-mode"; ORB_MR_SELF=1 timeout 300 python tools/build_once.py 27 20 3 2>&1 | tail -1
ORB_MR_SELF=1 ORB_PROFILE=1 ORB_DEBUG_SELECT=1 timeout 300 python tools/build_once.py 27 20 2 > $out/${tag}_2720_self_levels.txt 2>&1
