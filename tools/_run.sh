python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_c4.json 2> gpurun_out/r2k_c4.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_c3.json 2> gpurun_out/r2k_c3.err
