# parity tests + throughput of the example scenes given as arguments
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for s in "$@"; do MRT_JIT=${JIT:-1} python tools/bench_scenes.py --only $s; done
