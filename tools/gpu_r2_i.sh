set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_ipc.py tests/test_native_host.py -m gpu -q -k "group or ipc or multi_gpu" > gpurun_out/r2i_pytest_multi.txt 2>&1; tail -5 gpurun_out/r2i_pytest_multi.txt
timeout 300 python tools/host_overheads.py > gpurun_out/r2i_host_overheads.json 2>&1; cat gpurun_out/r2i_host_overheads.json
timeout 600 python bench.py --gpus 2 --no-cpu-baseline --steps 3 > gpurun_out/r2i_bench_group2.json 2> gpurun_out/r2i_bench_group2.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/r2i_bench_group2.json') if l.startswith('{')][-1]);print('group2 value',round(d['value']),'e2e',round(d['e2e']['value']),d['ms_per_step'],d['e2e']['ms_per_step'],json.dumps(d.get('cold_e2e')))"
