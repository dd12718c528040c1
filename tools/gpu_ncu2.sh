bash tools/gpu_variants.sh
bash tools/gpu_ncu.sh $1
