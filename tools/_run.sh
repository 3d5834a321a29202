python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long" 2>&1 | tail -15
timeout 600 python tools/long_waves.py 7 2>&1 | tail -4
