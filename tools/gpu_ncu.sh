# one ncu --set full capture of the path kernel (128 spp launch) after a plain run exited 0
set -x
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --spp 128 --no-cpu-baseline > gpurun_out/ncu_plain.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:path_kernel -c 1 -f -o gpurun_out/${1:-prof} python bench.py --steps 1 --warmup 3 --spp 128 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
