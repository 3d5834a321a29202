python tools/small_chunk_breakdown.py 20 2>&1 | tail -2
