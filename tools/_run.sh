timeout 300 python tools/enc_time.py 153391 3500 4 2000 20 2>&1 | tail -1
timeout 300 python tools/enc_time.py 153391 3500 4 2000 20 2>&1 | tail -1
timeout 300 python tools/enc_time.py 76695 7000 8 2000 20 2>&1 | tail -1
