set -x
mkdir -p gpurun_out
MRT_JIT=2 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
MRT_JIT_DUMP=$PWD/gpurun_out/jit_cb2.cubin MRT_JIT_DUMP_HEADER=$PWD/gpurun_out/jit_cb2.h python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_jit.json 2> gpurun_out/bench_jit.err; tail -5 gpurun_out/bench_jit.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_jit.json').read().strip().splitlines()[-1])
print("JIT VALUE", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "frac", round(l["roofline"]["frac"],4), l.get("jit"), l["clocks"])
PY
MRT_JIT=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nojit.json 2> gpurun_out/bench_nojit.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_nojit.json').read().strip().splitlines()[-1])
print("NOJIT VALUE", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "frac", round(l["roofline"]["frac"],4), l.get("jit"))
PY
