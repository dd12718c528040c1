# 16-byte sphere records in the BVH leaves: suite + Instance (the scene with sphere leaves) + ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2x_pytest_gpu.txt
for i in 1 2; do timeout 300 python tools/bench_scenes.py --only Instance 2>&1 | cut -c1-125; done
