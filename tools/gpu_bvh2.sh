for s in Minecraft CornellBox2 dof; do
MRT_JIT=0 python tools/bench_scenes.py --only $s
MRT_JIT=0 MRT_BVH_MIN=4 python tools/bench_scenes.py --only $s
done
