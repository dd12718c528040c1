# Round 2 experiment: t1 >= 0 folded into the hit update's minimum on the FMA pipe (MRT_T1_SIGN_ON_FMA_PIPE)
mkdir -p gpurun_out
bash tools/gpu_defines.sh "" "-DMRT_T1_SIGN_ON_FMA_PIPE=1" "" "-DMRT_T1_SIGN_ON_FMA_PIPE=1" 2>&1 | tee gpurun_out/r2q_t1sign.txt
MRT_JIT_DEFINES="-DMRT_T1_SIGN_ON_FMA_PIPE=1" timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jit and (primary_hits or shared_rng or fuzz or cluster)" 2>&1 | tail -3
