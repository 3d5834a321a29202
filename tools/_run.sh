for M in 16 64; do
DRICE_DEBUG=1 timeout 300 python tools/enc_time.py 76695 7000 $M 2000 10 2>&1 | grep -E "encode_tile|median" | tail -2 | cut -c1-150
done
DRICE_DEBUG=1 timeout 300 python tools/enc_time.py 153391 3500 64 2000 10 2>&1 | grep -E "encode_tile|median" | tail -2 | cut -c1-150
DRICE_ENC_LUT=0 timeout 300 python tools/enc_time.py 153391 3500 64 2000 10 2>&1 | grep -E "encode_tile|median" | tail -1 | cut -c1-150
