# Round 2, run E (1 GPU): statistical parity, precision ablation, compute-sanitizer, register-budget / smem-stack A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_statistics.py -m gpu -q > gpurun_out/r2e_pytest_stats.txt 2>&1; tail -40 gpurun_out/r2e_pytest_stats.txt
timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -q -k "wave or pooled" > gpurun_out/r2e_pytest_wave.txt 2>&1; tail -5 gpurun_out/r2e_pytest_wave.txt
timeout 600 python tools/precision_ablation.py gpurun_out/r2e_precision_default.json > gpurun_out/r2e_precision_default.txt 2>&1; tail -16 gpurun_out/r2e_precision_default.txt | cut -c1-330
MRT_LIB=$PWD/scratch/libmrt_precise.so MRT_JIT_CACHE=$PWD/gpurun_out/jitcache_precise timeout 600 python tools/precision_ablation.py gpurun_out/r2e_precision_precise.json > gpurun_out/r2e_precision_precise.txt 2>&1; tail -16 gpurun_out/r2e_precision_precise.txt | cut -c1-330
rm -rf gpurun_out/jitcache_precise
run() { tag=$1; shift; env "$@" timeout 300 python tools/bench_scenes.py --passes 128 --only3 > gpurun_out/r2e_scenes_$tag.jsonl 2> gpurun_out/r2e_scenes_$tag.err; echo "== $tag"; cut -c1-150 gpurun_out/r2e_scenes_$tag.jsonl; tail -2 gpurun_out/r2e_scenes_$tag.err; }
run mb8 MRT_X=1
run mb6 MRT_JIT_MINBLOCKS=6
run mb5 MRT_JIT_MINBLOCKS=5
run smem24_mb8 MRT_JIT_DEFINES=-DMRT_SMEM_STACK=24
run smem24_mb6 MRT_JIT_DEFINES=-DMRT_SMEM_STACK=24 MRT_JIT_MINBLOCKS=6
bash tools/gpu_sanitize.sh
