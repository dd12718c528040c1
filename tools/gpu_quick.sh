# quick GPU check: parity tests + one bench line (no CPU baseline)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -12 gpurun_out/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("VALUE", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "frac", round(l["roofline"]["frac"],4), "launch_ms", round(l["roofline"]["launch_ms"],2), l["clocks"])
PY
