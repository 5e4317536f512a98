python -m pytest tests -m gpu -q 2>&1 | tail -3
for s in 3 4 8; do echo "S=$s"; B=12 H=192 W=640 S=$s python tools/prof_step.py 30 2>&1 | tail -1; done
python tools/sweep_c5.py 12 > gpurun_out/r02_c5_sweep.txt 2> gpurun_out/c5.err; tail -3 gpurun_out/c5.err
B=8 H=320 W=1024 S=3 python tools/prof_step.py 5 > /dev/null 2>&1 && B=8 H=320 W=1024 S=3 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 11 --csv --log-file gpurun_out/launches_c3_r02.csv python tools/prof_step.py 5 > gpurun_out/ncu_c3.log 2>&1
B=12 H=192 W=640 S=4 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 11 --csv --log-file gpurun_out/launches_s4_r02.csv python tools/prof_step.py 5 > gpurun_out/ncu_s4.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json
