python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2h_smoke.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_c2.json 2> gpurun_out/r2h_c2.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2h_c2_ref.json 2> gpurun_out/r2h_c2_ref.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_c3.json 2> gpurun_out/r2h_c3.err
timeout 600 python tools/long_waves.py 7 2>&1 | tail -4 > gpurun_out/r2h_long.log
timeout 300 python tools/dbg_h5z_latency.py 2>&1 | tail -3 >> gpurun_out/r2h_long.log
