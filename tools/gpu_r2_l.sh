# Round 2: 256-bit loads in the BVH kernels (A/B against MRT_NO_LD256), suite
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2l_pytest_gpu.txt
for i in 1 2; do
echo "LD256"; timeout 300 python tools/bench_scenes.py 2>&1 | cut -c1-125 | tee -a gpurun_out/r2l_ld256.jsonl
echo "NO_LD256"; MRT_JIT_DEFINES="-DMRT_NO_LD256" timeout 300 python tools/bench_scenes.py 2>&1 | cut -c1-125 | tee -a gpurun_out/r2l_no_ld256.jsonl
done
