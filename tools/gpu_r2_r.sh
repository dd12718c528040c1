# bench.py with a COLD kernel cache and a single warm-up step: the timed region must still run on the specialised kernel
# (N = 1 and torchrun N = 2, where the wait has to be taken by both ranks together)
mkdir -p gpurun_out
MRT_JIT_CACHE=/tmp/mrt_cold1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2r_cold_n1.json 2> gpurun_out/r2r_cold_n1.err
MRT_JIT_CACHE=/tmp/mrt_cold2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 2 --warmup 1 --no-cold > gpurun_out/r2r_cold_n2.json 2> gpurun_out/r2r_cold_n2.err
MRT_JIT_CACHE=/tmp/mrt_cold3 timeout 600 python bench.py --gpus 2 --steps 2 --warmup 1 --no-cold > gpurun_out/r2r_cold_group2.json 2> gpurun_out/r2r_cold_group2.err
for f in n1 n2 group2; do python -c "import json;d=json.loads([l for l in open('gpurun_out/r2r_cold_$f.json') if l.startswith('{')][-1]);print('$f','value',round(d['value']),'e2e',round(d['e2e']['value']),d['roofline']['kernel'],d['jit'])" || tail -5 gpurun_out/r2r_cold_$f.err; done
