python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "straddles or long" 2>&1 | tail -15
python tools/small_chunk_breakdown.py 20 2>&1 | grep -E "decode|kernel times" | cut -c1-200
python tools/small_chunk_breakdown.py 60 2>&1 | grep -E "device-resident decode|kernel times" | cut -c1-200
