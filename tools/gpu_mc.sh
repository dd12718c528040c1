for mb in 3 4 6 8; do echo "MINBLOCKS $mb"; MRT_JIT=2 MRT_JIT_MINBLOCKS=$mb MRT_JIT_CACHE=off python tools/bench_scenes.py --only Minecraft; done
