# Round 2: ncu of the BVH kernels after the 256-bit loads (is the L1 data pipe still the bound?) + the rest of the suite
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r2n_pytest_gpu.txt
for sc in Instance:16 Minecraft:4 Mesh:32; do
  name=${sc%%:*}; passes=${sc##*:}
  MRT_JIT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r2n_${name} python tools/bench_scenes.py --only $name --passes $passes > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log | cut -c1-150
done
