python -m pytest tests -m gpu -x -q 2>&1 | tail -6
MRT_JIT=1 python tools/bench_scenes.py --only Instance
MRT_NO_BVH=1 python tools/bench_scenes.py --only Instance
