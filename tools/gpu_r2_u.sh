# last check of the round's final tree: full GPU suite + smoke + one short bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2u_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/r2u_bench.json') if l.startswith('{')][-1]);print('value',round(d['value']),'e2e',round(d['e2e']['value']),d['roofline']['kernel'],round(d['roofline']['frac'],4),d['clocks'])" || tail -5 gpurun_out/r2u_bench.err
