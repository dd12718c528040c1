# Round 2, pinhole entry point (cached first hit, rotated loop): suite, headline bench, every config
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2j_pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print("VALUE", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "frac", round(l["roofline"]["frac"],4), "launch_ms", round(l["roofline"]["launch_ms"],2), l["clocks"], l["image_mean_u8"])
PY
timeout 600 python tools/bench_scenes.py > gpurun_out/r2j_scenes.jsonl 2> gpurun_out/r2j_scenes.err; cut -c1-220 gpurun_out/r2j_scenes.jsonl
