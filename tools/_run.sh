timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not geometries and not front_end" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['all_kernels'])"
