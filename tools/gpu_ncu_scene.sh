# ncu --set full of the path kernel on one example scene: bash tools/gpu_ncu_scene.sh Mesh prof_name [jit-mode]
mkdir -p gpurun_out
MRT_JIT=${3:-0} python tools/bench_scenes.py --only $1 || exit 1
MRT_JIT=${3:-0} ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/$2 python tools/bench_scenes.py --only $1 > gpurun_out/ncu_scene.log 2>&1
tail -2 gpurun_out/ncu_scene.log
