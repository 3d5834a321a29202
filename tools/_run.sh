for h in 0 1; do echo -n "heavy=$h "; DRICE_DEC_HEAVY=$h timeout 300 python tools/_qb2.py 2>&1 | tail -1; done
timeout 300 python tools/quick_bench.py 153391 3500 4 2000 20 2>&1 | grep decode
timeout 300 python tools/small_chunk_breakdown.py 2000 2>&1 | grep -E "H5Z|device-resident decode"
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
