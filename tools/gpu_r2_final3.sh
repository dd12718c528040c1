# the driver's default bench line at the round's final tree (refreshes profiles/r2_bench_n1.json)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-200 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
