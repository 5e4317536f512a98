timeout 300 python -m pytest tests/test_disp_head.py tests/test_install_reference.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/disp_head_bench.py 2>&1 | tail -5 > gpurun_out/r02_disp_head.txt; cat gpurun_out/r02_disp_head.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:"disp_head" -s 64 -c 12 --csv --log-file gpurun_out/disp_head_launches_bwd.csv python tools/disp_head_bench.py > gpurun_out/ncu_dh2.log 2>&1
for k in disp_head_fwd_staged_kernel disp_head_gw_staged_kernel; do
  timeout 300 ncu --set full --clock-control none --kernel-name-base function -k $k -s 2 -c 1 -f -o /tmp/dh_$k python tools/disp_head_bench.py > gpurun_out/ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/dh_$k.ncu-rep "round 2, disparity head scale 0 (B=12, C=16, 192x640), ncu --set full --clock-control none" > gpurun_out/r02_${k}_ncu_full.txt
done
