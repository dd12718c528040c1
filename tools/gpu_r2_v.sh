# register budget of the small unrolled scenes: 10 resident blocks (48 registers, a few spills) against the default (55 -> 9 blocks)
mkdir -p gpurun_out
for sc in CornellBox Default dof CornellBox2; do
  for mb in "" 10; do
    echo "MINBLOCKS [$mb] $sc: $(MRT_JIT_MINBLOCKS=$mb timeout 300 python tools/bench_scenes.py --only $sc 2>&1 | python -c "import sys,json; print(round(json.loads(sys.stdin.readline())['gpu_mpaths_s']))")"
  done
done 2>&1 | tee gpurun_out/r2v_minblocks_small.txt
