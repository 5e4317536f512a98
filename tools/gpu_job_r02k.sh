python -m pytest tests/test_disp_head.py tests/test_install_reference.py -m gpu -q 2>&1 | tail -3
python tools/disp_head_bench.py 2>&1 | tail -6
