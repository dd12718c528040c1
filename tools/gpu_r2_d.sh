# Round 2, run D (1 GPU): wave kernel correctness + A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -k "wave or pooled" > gpurun_out/r2d_pytest.txt 2>&1; tail -30 gpurun_out/r2d_pytest.txt
run() { tag=$1; shift; env "$@" timeout 300 python tools/bench_scenes.py --passes 128 --only3 > gpurun_out/r2d_scenes_$tag.jsonl 2> gpurun_out/r2d_scenes_$tag.err; echo "== $tag"; cut -c1-200 gpurun_out/r2d_scenes_$tag.jsonl; tail -2 gpurun_out/r2d_scenes_$tag.err; }
run mega MRT_WAVE=0
run wave MRT_WAVE=1
run wave_mb5 MRT_WAVE=1 MRT_JIT_MINBLOCKS=5
run wave_mb6 MRT_WAVE=1 MRT_JIT_MINBLOCKS=6
run wave_s128 MRT_WAVE=1 MRT_JIT_DEFINES=-DMRT_WAVE_SLOTS=128
run wave_s32 MRT_WAVE=1 MRT_JIT_DEFINES=-DMRT_WAVE_SLOTS=32
run wave_starve16 MRT_WAVE=1 MRT_JIT_DEFINES=-DMRT_WAVE_STARVE=16
run wave_starve2 MRT_WAVE=1 MRT_JIT_DEFINES=-DMRT_WAVE_STARVE=2
for sc in Mesh:32 Instance:16 Minecraft:4; do
  name=${sc%%:*}; passes=${sc##*:}
  MRT_WAVE=1 MRT_JIT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r2d_${name}_wave python tools/bench_scenes.py --only $name --passes $passes > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
done
