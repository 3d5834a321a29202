for r in 20 2000 4000 5000; do DRICE_DEBUG=1 python tools/enc_time.py $r 7000 8 2000 50 2>&1 | grep -E "encode_tile|median" | tail -2 | cut -c1-110; done
python tools/small_chunk_breakdown.py 20 2>&1 | tail -8 | cut -c1-160
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
