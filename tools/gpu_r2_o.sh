# Round 2: 48-byte BVH nodes (fp16 half extents rounded up): suite + the three BVH scenes
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r2o_pytest_gpu.txt
for i in 1 2; do timeout 300 python tools/bench_scenes.py --only3 2>&1 | cut -c1-125 | tee -a gpurun_out/r2o_scenes.jsonl; done
