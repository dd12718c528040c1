set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_statistics.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2g_pytest.txt 2>&1; tail -12 gpurun_out/r2g_pytest.txt
timeout 600 python tools/precision_ablation.py gpurun_out/r2g_precision_default.json > gpurun_out/r2g_precision_default.txt 2>&1; grep "Instance\|CornellBox/" gpurun_out/r2g_precision_default.txt | cut -c1-330
timeout 300 python bench.py --no-cpu-baseline --no-configs --no-cold --steps 3 > gpurun_out/r2g_bench_refine.json 2> gpurun_out/r2g_bench_refine.err; cut -c1-160 gpurun_out/r2g_bench_refine.json
MRT_JIT_MINBLOCKS=10 timeout 300 python bench.py --no-cpu-baseline --no-configs --no-cold --steps 3 > gpurun_out/r2g_bench_refine_mb10.json 2> gpurun_out/r2g_bench_refine_mb10.err; cut -c1-160 gpurun_out/r2g_bench_refine_mb10.json
timeout 300 python tools/bench_scenes.py --passes 128 > gpurun_out/r2g_scenes.jsonl 2> gpurun_out/r2g_scenes.err; cut -c1-150 gpurun_out/r2g_scenes.jsonl
