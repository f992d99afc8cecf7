#!/bin/bash
# Final 1-GPU evidence of a round: the default bench line (C3 + legs C2, C4g, C4p), the ncu launch list of the same
# command, DRAM traffic per launch of one build of every config, --set full of one C3 build (raw page as CSV).
# usage (under gpurun): bash tools/gpu_session_final_1gpu.sh <tag>
tag=${1:-r02q}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; tail -3 $out/${tag}_bench.err | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --legs '' --no-cpu-baseline > $out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for cfg in "c3 27 16 uniform" "c2 24 12 uniform" "c4g 26 14 gaussian" "c4p 26 14 plummer"; do
  set -- $cfg
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file $out/${tag}_traffic_$1.csv python tools/build_once.py $2 $3 0 $4 > $out/${tag}_traffic_$1.log 2>&1
  echo "traffic $1 rc=$?"
done
timeout 900 ncu --set full --clock-control none -k regex:'k_partition_coop|k_partition_cells|k_sel_stream|k_sel_percell|k_sel_finish|k_sel_resolve|k_split' -c 80 -o /tmp/${tag}_full_c3 -f python tools/build_once.py 27 16 0 > $out/${tag}_full_c3.log 2>&1
echo "full c3 rc=$?"
ncu -i /tmp/${tag}_full_c3.ncu-rep --page raw --csv > $out/${tag}_full_c3_raw.csv 2>/dev/null
ls -la $out/${tag}_full_c3_raw.csv
