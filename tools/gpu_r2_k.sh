# Round 2: pinhole entry point + BVH micro-changes (branch-free lexicographic update, 8-byte stack entries, 24-bit leaf index):
# suite, every config (default aperture and aperture 0), register-budget A/B of the BVH kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2k_pytest_gpu.txt
timeout 600 python tools/bench_scenes.py > gpurun_out/r2k_scenes.jsonl 2> gpurun_out/r2k_scenes.err; cut -c1-120 gpurun_out/r2k_scenes.jsonl
timeout 600 python tools/bench_scenes.py --pinhole > gpurun_out/r2k_scenes_pinhole.jsonl 2> gpurun_out/r2k_scenes_pinhole.err; cut -c1-130 gpurun_out/r2k_scenes_pinhole.jsonl
for mb in 7 9 10; do
  echo "MINBLOCKS $mb"
  MRT_JIT_MINBLOCKS=$mb timeout 300 python tools/bench_scenes.py --only3 2>&1 | cut -c1-120 | tee -a gpurun_out/r2k_minblocks_$mb.jsonl
done
