mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pinhole" 2>&1 | tail -6 | cut -c1-200
