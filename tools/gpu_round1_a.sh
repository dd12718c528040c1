set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
./tools/microbench > gpurun_out/microbench.txt 2>&1; cat gpurun_out/microbench.txt
python tools/golden_gpu.py > gpurun_out/golden_gpu.txt 2>&1; cat gpurun_out/golden_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt
for spl in 128 256 1024; do MRT_SPP_PER_LAUNCH=$spl python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_spl$spl.json 2> gpurun_out/bench_spl$spl.err; cat gpurun_out/bench_spl$spl.json | cut -c1-400; done
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:path_kernel -c 1 -o gpurun_out/prof_r1c python bench.py --steps 1 --warmup 3 --spp 128 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
