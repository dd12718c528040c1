# Round-1 evidence run at the final build (1 GPU): what the driver runs + the ncu captures committed under profiles/
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r1b_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r1b_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.txt 2>&1; tail -2 gpurun_out/r1b_smoke.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b_bench_reference.json 2> gpurun_out/r1b_bench_reference.err; cut -c1-300 gpurun_out/r1b_bench_reference.json
python bench.py > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; cat gpurun_out/r1b_bench.json
MRT_JIT=0 python bench.py --no-cpu-baseline > gpurun_out/r1b_bench_generic.json 2> gpurun_out/r1b_bench_generic.err; cut -c1-200 gpurun_out/r1b_bench_generic.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:path_kernel_jit -c 1 -f -o gpurun_out/r1b_path_kernel_jit python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1b_ncu_full.log 2>&1
tail -2 gpurun_out/r1b_ncu_full.log
python tools/bench_scenes.py --cpu > gpurun_out/r1b_scenes.jsonl 2> gpurun_out/r1b_scenes.err; cat gpurun_out/r1b_scenes.jsonl
MRT_JIT=2 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r1b_mesh python tools/bench_scenes.py --only Mesh --passes 32 > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r1b_instance python tools/bench_scenes.py --only Instance --passes 16 > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log
