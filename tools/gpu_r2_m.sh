# Round 2: suite + bench after the NVRTC selection fix (256-bit loads need NVRTC >= 12.9; a torch process has 12.8 loaded)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2m_pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cold > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print("VALUE", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "frac", round(l["roofline"]["frac"],4), "launch_ms", round(l["roofline"]["launch_ms"],2), l["clocks"], l["jit"])
for c in l.get("configs", []): print(c["workload"], round(c["mpaths_s"]), c["jit"])
PY
