mkdir -p gpurun_out
MRT_JIT_CACHE=/tmp/mrt_cold9 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2w_cold_n1.json 2> gpurun_out/r2w_cold_n1.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/r2w_cold_n1.json') if l.startswith('{')][-1]);print('value',round(d['value']),'e2e',round(d['e2e']['value']),d['roofline']['kernel'],d['jit'])" || tail -5 gpurun_out/r2w_cold_n1.err
MRT_JIT=0 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2w_nojit_n1.json 2> gpurun_out/r2w_nojit_n1.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/r2w_nojit_n1.json') if l.startswith('{')][-1]);print('nojit value',round(d['value']),'e2e',round(d['e2e']['value']),d['roofline']['kernel'],d['jit'])" || tail -5 gpurun_out/r2w_nojit_n1.err
