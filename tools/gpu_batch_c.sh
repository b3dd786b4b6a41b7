#!/bin/bash
# ncu --set full captures of the kernels that carry the step (one launch each), exported to text by tools/ncu_export.py
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on --launch-count 1"
cap() {  # name, kernel regex, launch-skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$k --launch-skip $skip -f -o $O/r02_$name "$@" > $O/ncu_$name.log 2>&1
  python tools/ncu_export.py $O/r02_$name.ncu-rep > $O/r02_ncu_$name.txt 2>/dev/null
}
cap fps_l1        fps3_direct_kernel        3 python tools/probe_one.py fps_l1
cap bq_grid_l1    ball_query_grid_kernel    3 python tools/probe_one.py bq_l1
cap bq_build_l1   bq_grid_build_kernel      3 python tools/probe_one.py bq_l1
cap ffps_l2       ffps_cluster_kernel       3 python tools/probe_one.py ffps_l2
cap fused_l1s3    sa_fused_kernel           6 python tools/fused_probe.py 0
cap fused_l2s1    sa_fused_kernel           6 python tools/fused_probe.py 2
cap tc_hoist_l3   linear_tc_kernel          7 python tools/tc_probe.py 10
cap tc_l3_pooled  linear_tc_kernel          6 python tools/tc_probe.py 2
cap tc_l4_pooled  linear_tc_kernel          6 python tools/tc_probe.py 3
cap expand_l4     hoist_expand_split_kernel 3 python tools/probe_one.py expand_l4
cap nms           bev_nms_kernel            3 python tools/probe_one.py nms
grep -h -E "kernel:|gpu__time_duration|pipe_tensor_cycles|dram__bytes" $O/r02_ncu_*.txt | head -60
