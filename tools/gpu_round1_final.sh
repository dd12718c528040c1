# Round-1 evidence run (1 GPU): what the driver runs + the ncu captures committed under profiles/
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r1_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r1_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.txt 2>&1; tail -2 gpurun_out/r1_smoke.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err; cat gpurun_out/r1_bench_reference.json | cut -c1-300
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; cat gpurun_out/r1_bench.json
MRT_JIT=0 python bench.py --no-cpu-baseline > gpurun_out/r1_bench_generic.json 2> gpurun_out/r1_bench_generic.err; cat gpurun_out/r1_bench_generic.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:path_kernel_jit -c 1 -f -o gpurun_out/r1_path_kernel_jit python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
tail -2 gpurun_out/r1_ncu_full.log
python tools/bench_scenes.py --cpu > gpurun_out/r1_scenes.jsonl 2> gpurun_out/r1_scenes.err; cat gpurun_out/r1_scenes.jsonl
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -ftz=true -o tools/microbench tools/microbench.cu && ./tools/microbench > gpurun_out/r1_microbench.txt 2>&1
