python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -4
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "eager", d["e2e_eager"]["ms_per_step"], "fp32", d["e2e_fp32_pyramid"]["ms_per_step"], "c3", d["c3"]["ms_per_step"], "c4", d["c4"]["ms_per_step"], "ddp", d["ddp_step"]["ms_per_step"], "roof", d["roofline"]["frac"], d["roofline_issue"]["frac"], d["clocks"])
PY
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
python tools/disp_head_bench.py 2>&1 | tail -5 > gpurun_out/r02_disp_head.txt; cat gpurun_out/r02_disp_head.txt
