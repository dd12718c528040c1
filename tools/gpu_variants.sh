# bench every scratch/libmrt_*.so variant (kernel experiments) with the generic kernel (MRT_JIT=0)
# and the in-tree library with the run-time specialised kernel; prints VALUE per variant
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || { echo "FAILED $name"; tail -3 gpurun_out/v.err; return; }
  python - "$name" <<'PY'
import json,sys
l=json.loads(open('gpurun_out/v.json').read().strip().splitlines()[-1])
print("VARIANT", sys.argv[1], "VALUE", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "mean", round(l["image_mean_u8"],3), l["clocks"]["sm_mhz"], l["clocks"]["reasons"], "jit_launches", l["jit"]["launches"])
PY
}
for so in scratch/libmrt_*.so; do run "$so generic" MRT_LIB=$PWD/$so MRT_JIT=0; done
run "in-tree jit" MRT_JIT=1
