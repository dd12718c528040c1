# Round 2 evidence run at the final build (1 GPU): suite, smoke, bench (+ reference arm), ncu launch list, ncu --set full of the
# headline kernel, of its pinhole entry point and of the three BVH scenes, per-scene throughput (default aperture and aperture 0),
# generated headers + cubins
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-160 gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; cut -c1-160 gpurun_out/r2_bench_reference_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2_ncu_launch.log 2>&1; tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-200
MRT_JIT=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:path_kernel_jit -s 1 -c 1 -f -o gpurun_out/r2_path_kernel_jit python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs --no-cold > gpurun_out/r2_ncu_full.log 2>&1; tail -1 gpurun_out/r2_ncu_full.log | cut -c1-200
MRT_JIT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_kernel_jit_pinhole -s 1 -c 1 -f -o gpurun_out/r2_CornellBox2_pinhole python tools/bench_scenes.py --only CornellBox2 --passes 128 --pinhole > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log | cut -c1-150
for sc in Mesh:32 Instance:16 Minecraft:4; do
  name=${sc%%:*}; passes=${sc##*:}
  MRT_JIT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:path_kernel -s 1 -c 1 -f -o gpurun_out/r2_${name} python tools/bench_scenes.py --only $name --passes $passes > gpurun_out/ncu_scene.log 2>&1; tail -1 gpurun_out/ncu_scene.log | cut -c1-150
done
timeout 600 python tools/bench_scenes.py --cpu > gpurun_out/r2_scenes.jsonl 2> gpurun_out/r2_scenes.err; cut -c1-200 gpurun_out/r2_scenes.jsonl
timeout 600 python tools/bench_scenes.py --pinhole > gpurun_out/r2_scenes_pinhole.jsonl 2> gpurun_out/r2_scenes_pinhole.err; cut -c1-200 gpurun_out/r2_scenes_pinhole.jsonl
bash tools/gpu_checked.sh
python tools/dump_jit.py gpurun_out/jit_r2 > gpurun_out/r2_dump.log 2>&1; tail -3 gpurun_out/r2_dump.log
