python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --no-e2e --steps 30 --warmup 5 > gpurun_out/bench_pre.json 2> gpurun_out/bench_pre.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_pre.json").read().strip().splitlines()[-1])
print("headline ms", d["ms_per_step"], "c3", d.get("c3", {}).get("ms_per_step"), "c4", d.get("c4", {}).get("ms_per_step"))
PY
for s in 1 3 4 5 8; do echo "S=$s"; B=12 H=192 W=640 S=$s python tools/prof_step.py 30 2>&1 | tail -1; done
B=8 H=320 W=1024 S=3 python tools/prof_step.py 5 > /dev/null 2>&1 && B=8 H=320 W=1024 S=3 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 11 --csv --log-file gpurun_out/launches_c3_pre.csv python tools/prof_step.py 5 > gpurun_out/ncu_c3.log 2>&1
tail -1 gpurun_out/ncu_c3.log
