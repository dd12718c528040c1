# bench the specialised kernel under several MRT_JIT_DEFINES settings (kernel experiments); prints VALUE per variant
# usage: bash tools/gpu_defines.sh "<defines A>" "<defines B>" ...   ("" = the default build)
mkdir -p gpurun_out
for defs in "$@"; do
  MRT_JIT_DEFINES="$defs" MRT_JIT=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || { echo "FAILED [$defs]"; tail -3 gpurun_out/v.err; continue; }
  python - "$defs" <<'PY'
import json,sys
l=json.loads(open('gpurun_out/v.json').read().strip().splitlines()[-1])
print("DEFINES [%s]" % sys.argv[1], "VALUE", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "mean", round(l["image_mean_u8"],4), l["clocks"]["sm_mhz"], l["clocks"]["reasons"], "jit_launches", l["jit"]["launches"])
PY
done
