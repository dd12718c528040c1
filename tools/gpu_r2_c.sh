# Round 2, run C (2 GPUs): device-group tests, native host --gpus 2, bench through one group context and under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_native_host.py -m gpu -q > gpurun_out/r2c_pytest_gpu.txt 2>&1; tail -25 gpurun_out/r2c_pytest_gpu.txt
timeout 600 python bench.py --gpus 2 --no-cpu-baseline --steps 3 > gpurun_out/r2c_bench_group2.json 2> gpurun_out/r2c_bench_group2.err; cut -c1-1200 gpurun_out/r2c_bench_group2.json; tail -3 gpurun_out/r2c_bench_group2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 > gpurun_out/r2c_bench_torchrun2.json 2> gpurun_out/r2c_bench_torchrun2.err; cut -c1-1200 gpurun_out/r2c_bench_torchrun2.json; tail -3 gpurun_out/r2c_bench_torchrun2.err
