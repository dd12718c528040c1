# A/B of the pixel->lane mapping (MRT_TILE=1: 8x4 tile per warp, 0: 32 pixels of a row): headline bench + every scene
mkdir -p gpurun_out
for t in 1 0; do
  MRT_TILE=$t python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || { echo FAILED; tail -3 gpurun_out/v.err; }
  python - $t <<'PY'
import json,sys
l=json.loads(open('gpurun_out/v.json').read().strip().splitlines()[-1])
print("TILE", sys.argv[1], "VALUE", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "mean", round(l["image_mean_u8"],4), l["clocks"]["reasons"])
PY
  MRT_TILE=$t python tools/bench_scenes.py | cut -c1-135
done
