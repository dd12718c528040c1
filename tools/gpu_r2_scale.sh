# Round 2 scaling run on one 8-GPU box: multi-GPU tests, torchrun bench at N = 2, 4, 8 (IPC band gather; N = 8 also with the
# round-1 NCCL reduce), bench through ONE group context at N = 8, the one-shot binary on 8 GPUs
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_ipc.py tests/test_native_host.py -m gpu -q -k "group or ipc or multi_gpu" > gpurun_out/r2s_pytest_multi.txt 2>&1; tail -6 gpurun_out/r2s_pytest_multi.txt
port=29600
for n in 2 4 8; do
  port=$((port+1))
  extra="--no-cold"; [ $n = 8 ] && extra=""
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n $extra > gpurun_out/r2s_scale_n$n.json 2> gpurun_out/r2s_scale_n$n.err
  python -c "import json;d=json.load(open('gpurun_out/r2s_scale_n$n.json'));print($n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'launch_ms',d['roofline']['launch_ms'],'ms/step',d['ms_per_step'],d['e2e']['ms_per_step'],d.get('cold_e2e'))" || tail -5 gpurun_out/r2s_scale_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --no-cold --nccl-reduce > gpurun_out/r2s_scale_n8_nccl.json 2> gpurun_out/r2s_scale_n8_nccl.err
python -c "import json;d=json.load(open('gpurun_out/r2s_scale_n8_nccl.json'));print('8 nccl value',round(d['value']),'e2e',round(d['e2e']['value']),d['ms_per_step'],d['e2e']['ms_per_step'])" || tail -5 gpurun_out/r2s_scale_n8_nccl.err
timeout 900 python bench.py --gpus 8 --no-cold > gpurun_out/r2s_group_n8.json 2> gpurun_out/r2s_group_n8.err
python -c "import json;d=json.load(open('gpurun_out/r2s_group_n8.json'));print('8 group value',round(d['value']),'e2e',round(d['e2e']['value']),d['ms_per_step'],d['e2e']['ms_per_step'])" || tail -5 gpurun_out/r2s_group_n8.err
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2s_scale_n1.json 2> gpurun_out/r2s_scale_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2s_scale_n1.json'));print('1 value',round(d['value']),'e2e',round(d['e2e']['value']))"
