# torchrun N = 8 with a COLD kernel cache (every rank compiles; the wait for the specialised kernel is taken by all ranks together)
mkdir -p gpurun_out
MRT_JIT_CACHE=/tmp/mrt_cold8 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 5 --warmup 3 --no-cold > gpurun_out/r2y_cold_n8.json 2> gpurun_out/r2y_cold_n8.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/r2y_cold_n8.json') if l.startswith('{')][-1]);print('n8 cold value',round(d['value']),'e2e',round(d['e2e']['value']),d['roofline']['kernel'],d['jit'])" || tail -5 gpurun_out/r2y_cold_n8.err
