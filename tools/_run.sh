timeout 300 python tools/enc_time.py 153391 3500 256 2000 20 2>&1 | tail -1 | cut -c1-100
timeout 300 python tools/enc_time.py 76695 7000 1 2000 20 2>&1 | tail -1 | cut -c1-100
timeout 300 python tools/enc_time.py 76695 7000 64 2000 20 2>&1 | tail -1 | cut -c1-100
DRICE_ENC_LUT=0 timeout 300 python tools/enc_time.py 153391 3500 4 2000 20 2>&1 | tail -1 | cut -c1-100
