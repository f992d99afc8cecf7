out=gpurun_out; tag=r02x
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > $out/${tag}_gpu_tests.txt 2>&1; echo "gpu tests rc=$?"; tail -4 $out/${tag}_gpu_tests.txt | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python tools/bbox_once.py 27 16 > $out/${tag}_bbox.txt 2>&1; tail -5 $out/${tag}_bbox.txt
timeout 200 python bench.py --steps 5 --warmup 3 --legs '' --no-cpu-baseline > $out/${tag}_bench_c3_short.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
