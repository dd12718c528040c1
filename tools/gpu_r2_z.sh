mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 --no-cold --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2z_bench.json') if x.startswith('{')][-1])
print("VALUE", round(l["value"],1), "e2e", round(l["e2e"]["value"],1))
for c in l.get("configs", []): print(c["workload"], round(c["mpaths_s"]), round(c["aperture0_mpaths_s"]), c["jit"])
PY
tail -3 gpurun_out/r2z_bench.err
