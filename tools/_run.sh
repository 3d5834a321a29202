for i in 1 2; do
DRICE_ENC_WORKERS=12 timeout 300 python tools/enc_time.py 153391 3500 4 2000 20 2>&1 | tail -1 | cut -c1-80
timeout 300 python tools/enc_time.py 153391 3500 4 2000 20 2>&1 | tail -1 | cut -c1-80
done
