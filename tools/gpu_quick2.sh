# parity suite + one bench line each for the specialised and the generic kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -2 gpurun_out/pytest_gpu.txt
for j in 1 0; do
MRT_JIT=$j python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - $j <<'PY'
import json,sys
l=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("JIT" if sys.argv[1]=="1" else "GENERIC", "VALUE", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "frac", round(l["roofline"]["frac"],4), l["clocks"]["reasons"])
PY
done
MRT_JIT=1 python tools/bench_scenes.py --only Instance | cut -c1-130
