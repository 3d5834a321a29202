timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
DRICE_DEBUG=1 timeout 120 python tools/enc_time.py 153391 3500 4 2000 10 2>&1 | grep -E "encode_tile|lut=" | tail -3 | cut -c1-200
DRICE_DEBUG=1 timeout 120 python tools/enc_time.py 76696 7000 8 2000 10 2>&1 | grep -E "encode_tile|lut=" | tail -3 | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c2_v2.json 2> gpurun_out/r2_c2_v2.err
