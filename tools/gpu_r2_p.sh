# Round 2 experiment: thin-lens loop with pre-generated camera rays (MRT_RING = cadence) against the plain loop
mkdir -p gpurun_out
bash tools/gpu_defines.sh "" "-DMRT_RING=2" "-DMRT_RING=4" "-DMRT_RING=8" 2>&1 | tee gpurun_out/r2p_ring.txt
