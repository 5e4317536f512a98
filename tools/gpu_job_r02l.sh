python tools/disp_head_bench.py > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base function -k regex:disp_head -c 60 --csv --log-file gpurun_out/disp_head_launches.csv python tools/disp_head_bench.py > gpurun_out/ncu_dh.log 2>&1
for k in disp_head_fwd_tile_kernel disp_head_gx_tile_kernel disp_head_gw_kernel; do
  ncu --set full --clock-control none --kernel-name-base function -k $k -s 2 -c 1 -f -o /tmp/dh_$k python tools/disp_head_bench.py > gpurun_out/ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/dh_$k.ncu-rep "round 2, disparity head scale 0 (B=12, C=16, 192x640), ncu --set full --clock-control none" > gpurun_out/r02_${k}_ncu_full.txt
done
