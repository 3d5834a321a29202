python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_c2.json 2> gpurun_out/r2f_c2.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_c2_ref.json 2> gpurun_out/r2f_c2_ref.err
