# Round 2, run A (1 GPU): suite + smoke + bench at the ABI-v2 build; pooled-kernel A/B on the BVH scenes; fresh ncu captures
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.txt 2>&1; tail -15 gpurun_out/r2a_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.txt 2>&1; tail -2 gpurun_out/r2a_smoke.txt
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cut -c1-400 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
for pool in 0 1; do
  MRT_POOL=$pool timeout 300 python tools/bench_scenes.py --passes 128 > gpurun_out/r2a_scenes_pool$pool.jsonl 2> gpurun_out/r2a_scenes_pool$pool.err
  cut -c1-160 gpurun_out/r2a_scenes_pool$pool.jsonl
done
for sc in Mesh:32 Instance:16 Minecraft:4; do
  name=${sc%%:*}; passes=${sc##*:}
  MRT_JIT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r2a_${name}_pool python tools/bench_scenes.py --only $name --passes $passes > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
done
