python tools/fuzz_gpu.py 1500 11 2>&1 | tail -2
python tools/fuzz_gpu.py 1500 12 2>&1 | tail -2
DRICE_LOCATE_DIRECT=1 DRICE_DEC_SORT=2 python tools/fuzz_gpu.py 1000 13 2>&1 | tail -2
DRICE_ENC_WORKERS=8 DRICE_PARSE_WIDE=0 python tools/fuzz_gpu.py 800 14 2>&1 | tail -2
DRICE_ENC_WORKERS=24 DRICE_ENC_STAGE_WORDS=64 DRICE_DEC_SORT=2 python tools/fuzz_gpu.py 800 15 2>&1 | tail -2
DRICE_ENC_LUT=0 DRICE_LOCATE_SCAN=0 DRICE_PARSE_LONG_SHORT=0 python tools/fuzz_gpu.py 800 16 2>&1 | tail -2
