python -m pytest tests -m gpu -q 2>&1 | tail -4
for th in 0 64 112 160; do echo "PML_TH=$th"; PML_TH=$th B=8 H=320 W=1024 S=3 python tools/prof_step.py 30 2>&1 | tail -1; done
for th in 0 64 96; do echo "S5 PML_TH=$th"; PML_TH=$th B=12 H=192 W=640 S=5 python tools/prof_step.py 30 2>&1 | tail -1; done
B=8 H=320 W=1024 S=3 python tools/prof_step.py 5 > /dev/null 2>&1 && B=8 H=320 W=1024 S=3 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 11 --csv --log-file gpurun_out/launches_c3_pre.csv python tools/prof_step.py 5 > gpurun_out/ncu_c3.log 2>&1
tail -1 gpurun_out/ncu_c3.log
