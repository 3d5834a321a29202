timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_wave or errors or lane_parser" 2>&1 | tail -15
timeout 600 python tools/long_waves.py 5 2>&1 | tail -5
