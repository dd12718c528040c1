mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for mb in "" 8 9; do
  MRT_JIT_MINBLOCKS=$mb MRT_JIT_CACHE=off MRT_JIT_DUMP=$PWD/gpurun_out/jit_mb$mb.cubin python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || { echo FAILED $mb; tail -3 gpurun_out/v.err; continue; }
  python - "$mb" <<'PY'
import json,sys
l=json.loads(open('gpurun_out/v.json').read().strip().splitlines()[-1])
print("MINBLOCKS", sys.argv[1], "VALUE", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "jit_launches", l["jit"]["launches"])
PY
  cuobjdump -res-usage gpurun_out/jit_mb$mb.cubin | tail -1
done
