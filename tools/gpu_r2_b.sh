# Round 2, run B (1 GPU): full suite (no -x), flat BVH walk A/B against round 1's nested loops, register budgets, ncu
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.txt 2>&1; tail -25 gpurun_out/r2b_pytest_gpu.txt
run() { tag=$1; shift; env "$@" timeout 300 python tools/bench_scenes.py --passes 128 --only3 > gpurun_out/r2b_scenes_$tag.jsonl 2> gpurun_out/r2b_scenes_$tag.err; echo "== $tag"; cut -c1-150 gpurun_out/r2b_scenes_$tag.jsonl; tail -2 gpurun_out/r2b_scenes_$tag.err; }
run walk MRT_X=1
run v1 MRT_JIT_DEFINES=-DMRT_WALK_V1
run v1_unrolled_mesh MRT_JIT_DEFINES=-DMRT_WALK_V1 MRT_MESH_VIA_BVH=0
run walk_mb5 MRT_JIT_MINBLOCKS=5
run walk_mb6 MRT_JIT_MINBLOCKS=6
run walk_mb4 MRT_JIT_MINBLOCKS=4
for sc in Mesh:32 Instance:16 Minecraft:4; do
  name=${sc%%:*}; passes=${sc##*:}
  MRT_JIT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r2b_${name}_walk python tools/bench_scenes.py --only $name --passes $passes > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
done
python tools/dump_jit.py gpurun_out/jit_r2b > gpurun_out/r2b_dump.log 2>&1; tail -3 gpurun_out/r2b_dump.log
