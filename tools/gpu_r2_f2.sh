set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ipc.py -m gpu -q > gpurun_out/r2f2_pytest_ipc.txt 2>&1; tail -15 gpurun_out/r2f2_pytest_ipc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --no-cold > gpurun_out/r2f2_bench_torchrun2_ipc.json 2> gpurun_out/r2f2_bench_torchrun2_ipc.err; cut -c1-300 gpurun_out/r2f2_bench_torchrun2_ipc.json; grep -o '"e2e": {[^}]*}' gpurun_out/r2f2_bench_torchrun2_ipc.json | cut -c1-200; tail -3 gpurun_out/r2f2_bench_torchrun2_ipc.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --no-cold --nccl-reduce > gpurun_out/r2f2_bench_torchrun2_nccl.json 2> gpurun_out/r2f2_bench_torchrun2_nccl.err; cut -c1-300 gpurun_out/r2f2_bench_torchrun2_nccl.json; grep -o '"e2e": {[^}]*}' gpurun_out/r2f2_bench_torchrun2_nccl.json | cut -c1-200; tail -3 gpurun_out/r2f2_bench_torchrun2_nccl.err
timeout 300 python tools/diag_instance.py > gpurun_out/r2f2_diag.txt 2>&1; tail -6 gpurun_out/r2f2_diag.txt
