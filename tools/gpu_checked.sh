# The GPU test suite through the CHECKED build of the library (tools/build_variant.sh checked "-DMRT_CHECKED=1": every
# table index the kernels form is bounds-checked, a violation traps) — this pool's substitute for compute-sanitizer
# memcheck, which its GPUs do not admit.  Log -> gpurun_out/r2_checked_build_suite.txt
mkdir -p gpurun_out
MRT_LIB=$PWD/scratch/libmrt_checked.so timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_checked_build_suite.txt 2>&1
tail -6 gpurun_out/r2_checked_build_suite.txt
