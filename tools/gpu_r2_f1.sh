set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.txt 2>&1; tail -30 gpurun_out/r2f_pytest_gpu.txt
bash tools/gpu_checked.sh
