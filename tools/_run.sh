python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2i_c2.json 2> gpurun_out/r2i_c2.err
