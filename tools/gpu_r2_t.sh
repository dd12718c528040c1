# soak of the bit-identity / fuzz tests with the 48-byte BVH nodes: more seeds than the suite's defaults
mkdir -p gpurun_out
MRT_FUZZ_BVH_SEEDS=48 MRT_FUZZ_MESH_SEEDS=40 MRT_FUZZ_SEEDS=120 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "scene_bvh or triangle_bvh or fuzz or cluster" > gpurun_out/r2_soak.txt 2>&1
tail -5 gpurun_out/r2_soak.txt
