timeout 300 python tools/enc_time.py 153391 3500 4 2000 20 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/enc_time.py 76695 7000 8 2000 20 2>&1 | tail -1 | cut -c1-100
timeout 300 python tools/enc_time.py 153391 3500 256 2000 20 2>&1 | tail -1 | cut -c1-100
timeout 300 python tools/enc_time.py 2000 7000 8 2000 20 2>&1 | tail -1 | cut -c1-100
timeout 300 python tools/enc_time.py 20 7000 8 20 20 2>&1 | tail -1 | cut -c1-100
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
