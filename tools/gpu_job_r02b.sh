python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err; tail -2 gpurun_out/r02_bench_h.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-250 gpurun_out/r02_bench_ref.json
python tools/prof_step.py 3 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 24 --csv --log-file gpurun_out/launches_r02.csv python tools/prof_step.py 3 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
