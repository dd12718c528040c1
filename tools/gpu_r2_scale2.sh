# Round 2 scaling run at the final build on one 8-GPU box: multi-GPU tests, torchrun bench at N = 1, 2, 4, 8 as the driver
# launches it (one NCCL reduce of the films), N = 8 also with the IPC band gather and through ONE group context (+ one-shot binary)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_ipc.py tests/test_native_host.py -m gpu -q -k "group or ipc or multi_gpu" > gpurun_out/r2s_pytest_multi.txt 2>&1; tail -3 gpurun_out/r2s_pytest_multi.txt
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2s_scale_n1.json 2> gpurun_out/r2s_scale_n1.err
port=29700
for n in 2 4 8; do
  port=$((port+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --no-cold > gpurun_out/r2s_scale_n$n.json 2> gpurun_out/r2s_scale_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29750 bench.py --gpus 8 --no-cold --ipc-gather > gpurun_out/r2s_scale_n8_ipc.json 2> gpurun_out/r2s_scale_n8_ipc.err
timeout 900 python bench.py --gpus 8 > gpurun_out/r2s_group_n8.json 2> gpurun_out/r2s_group_n8.err
for f in n1 n2 n4 n8 n8_ipc; do python -c "import json;d=json.loads([l for l in open('gpurun_out/r2s_scale_$f.json') if l.startswith('{')][-1]);print('$f','value',round(d['value']),'e2e',round(d['e2e']['value']),'launch_ms',d['roofline']['launch_ms'],'ms/step',d['ms_per_step'],d['e2e']['ms_per_step'])" || tail -5 gpurun_out/r2s_scale_$f.err; done
python -c "import json;d=json.loads([l for l in open('gpurun_out/r2s_group_n8.json') if l.startswith('{')][-1]);print('8 group value',round(d['value']),'e2e',round(d['e2e']['value']),d['ms_per_step'],d['e2e']['ms_per_step'],json.dumps(d.get('cold_e2e')))" || tail -5 gpurun_out/r2s_group_n8.err
