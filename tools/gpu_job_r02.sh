python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -2 gpurun_out/r02_bench_e.err
python tools/sweep_c5.py 12 > gpurun_out/r02_c5_sweep.txt 2> gpurun_out/r02_c5.err
python tools/prof_step.py 5 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name-base function -k sweep_kernel -s 3 -c 1 -o gpurun_out/sweep_r02 python tools/prof_step.py 5 > gpurun_out/ncu_s.log 2>&1
python tools/prof_step.py 3 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_r02.csv python tools/prof_step.py 3 > gpurun_out/ncu_l.log 2>&1
tail -3 gpurun_out/ncu_s.log
