# ncu --set full of the secondary kernels (headline workload) and of the multi-frame kernels (BASELINE config 3);
# the reports stay on the box (17 MB each), only their summaries come back
python tools/prof_step.py 5 > gpurun_out/plain.log 2>&1 || exit 1
for k in prep_kernel finalize_image_kernel scale_grads_kernel disp_sum_kernel; do
  ncu --set full --clock-control none --kernel-name-base function -k $k -s 3 -c 1 -f -o /tmp/r02_$k python tools/prof_step.py 5 > gpurun_out/ncu_$k.log 2>&1
  python tools/ncu_summary.py /tmp/r02_$k.ncu-rep "round 2, headline workload (B=12, 192x640, S=2), ncu --set full --clock-control none" > gpurun_out/r02_${k}_ncu_full.txt
done
export B=8 H=320 W=1024 S=3
python tools/prof_step.py 5 > gpurun_out/plain_c3.log 2>&1 || exit 1
ncu --set full --clock-control none --kernel-name-base function -k select_prepass_kernel -s 3 -c 1 -f -o /tmp/r02_pp python tools/prof_step.py 5 > gpurun_out/ncu_pp.log 2>&1
python tools/ncu_summary.py /tmp/r02_pp.ncu-rep "round 2, BASELINE config 3 (B=8, 320x1024, S=3), ncu --set full --clock-control none" > gpurun_out/r02_select_prepass_ncu_full.txt
# sweep launches of one step in order: mode 1 (forward), mode 3 (lone frame, scalar), mode 2 (adjoint of the first pair)
ncu --set full --clock-control none --kernel-name-base function -k sweep_kernel -s 9 -c 3 -f -o /tmp/r02_sw python tools/prof_step.py 5 > gpurun_out/ncu_sw.log 2>&1
ncu -i /tmp/r02_sw.ncu-rep --page raw --csv > gpurun_out/r02_c3_sweeps_raw.csv
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
