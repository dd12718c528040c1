# Round-1 evidence refresh at the final build (1 GPU)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r1_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.txt 2>&1; tail -2 gpurun_out/r1_smoke.txt
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; cut -c1-250 gpurun_out/r1_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launch.log 2>&1
python tools/bench_scenes.py --cpu > gpurun_out/r1_scenes.jsonl 2> gpurun_out/r1_scenes.err; cut -c1-200 gpurun_out/r1_scenes.jsonl
MRT_JIT=2 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r1_mesh python tools/bench_scenes.py --only Mesh --passes 32 > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
MRT_JIT=2 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r1_instance python tools/bench_scenes.py --only Instance --passes 16 > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
MRT_JIT=2 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r1_minecraft python tools/bench_scenes.py --only Minecraft --passes 4 > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
