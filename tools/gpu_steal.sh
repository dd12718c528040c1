mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for spp in 1024 128; do
  timeout 300 python bench.py --steps 3 --warmup 3 --spp $spp --no-cpu-baseline > gpurun_out/v.json 2> gpurun_out/v.err || { echo FAILED; tail -3 gpurun_out/v.err; continue; }
  python - $spp <<'PY'
import json,sys
l=json.loads(open('gpurun_out/v.json').read().strip().splitlines()[-1])
print("SPP", sys.argv[1], "VALUE", round(l["value"],1), "frac", round(l["roofline"]["frac"],4), "mean", round(l["image_mean_u8"],3), "jit", l["jit"]["launches"])
PY
done
