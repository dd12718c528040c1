# bench every scratch/libmrt_*.so variant (kernel experiments); prints VALUE per variant
mkdir -p gpurun_out
for so in scratch/libmrt_*.so; do
  MRT_LIB=$PWD/$so python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || { echo "FAILED $so"; tail -3 gpurun_out/v.err; continue; }
  python - "$so" <<'PY'
import json,sys
l=json.loads(open('gpurun_out/v.json').read().strip().splitlines()[-1])
print("VARIANT", sys.argv[1], "VALUE", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "mean", round(l["image_mean_u8"],3), l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
PY
done
