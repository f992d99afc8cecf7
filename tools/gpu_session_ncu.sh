#!/bin/bash
# One 1-GPU box session: ncu evidence of one C3 / C2 / C4 build each (no warm-up build: ncu times kernels one by one).
#  (a) launch list + DRAM bytes of every launch (-> tools/ncu_traffic.py -> profiles/r02_ncu_traffic.json)
#  (b) --set full of the C3 build's kernels with source (-> tools/ncu_brief.py)
# usage (under gpurun): bash tools/gpu_session_ncu.sh <tag>
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for cfg in "c3 27 16 uniform" "c2 24 12 uniform" "c4g 26 14 gaussian" "c4p 26 14 plummer"; do
  set -- $cfg
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file $out/${tag}_traffic_$1.csv python tools/build_once.py $2 $3 0 $4 > $out/${tag}_traffic_$1.log 2>&1
  echo "traffic $1 rc=$?"
done
# (the .ncu-rep with sources is ~100 MB: it stays on the box; the raw page of every captured launch comes back as CSV)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_partition_coop|k_partition_cells|k_sel_stream|k_sel_percell|k_sel_finish|k_split' -c 70 -o /tmp/${tag}_full_c3 -f python tools/build_once.py 27 16 0 > $out/${tag}_full_c3.log 2>&1
echo "full c3 rc=$?"; ls -la /tmp/${tag}_full_c3.ncu-rep
ncu -i /tmp/${tag}_full_c3.ncu-rep --page raw --csv > $out/${tag}_full_c3_raw.csv 2>/dev/null
ls -la $out/${tag}_full_c3_raw.csv
