# strong-scaling run as the driver does it: N = 1,2,4,8 (or the list given) back to back
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  python - $N <<'PY'
import json,sys
try:
    l=json.loads(open(f'gpurun_out/scale_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print("SCALE N", l["n_gpus"], "VALUE", round(l["value"],1), "ms/step", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["value"],1), "launches", l["gpu_launches"], "jit", l["jit"]["launches"], l["clocks"]["sm_mhz"], l["clocks"]["reasons"])
except Exception as e:
    print("SCALE N", sys.argv[1], "FAILED", e); print(open(f'gpurun_out/scale_n{sys.argv[1]}.err').read()[-1500:])
PY
done
