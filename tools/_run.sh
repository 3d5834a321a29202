timeout 600 python tools/long_waves.py 7 2>&1 | tail -4
timeout 300 python tools/dbg_h5z_latency.py 2>&1 | tail -3
