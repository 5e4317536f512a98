timeout 300 python -m pytest tests/test_disp_head.py tests/test_install_reference.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/disp_head_bench.py 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:disp_head -c 14 --csv --log-file gpurun_out/disp_head_launches.csv python tools/disp_head_bench.py > gpurun_out/ncu_dh.log 2>&1
