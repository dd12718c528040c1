set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2h_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.txt 2>&1; tail -2 gpurun_out/r2h_smoke.txt
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; cut -c1-200 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cold", d.get("cold_e2e"))
for c in d.get("configs", []): print(c["workload"], round(c["mpaths_s"]), c.get("cpu_baseline", {}).get("value"))
PY
bash tools/gpu_checked.sh
