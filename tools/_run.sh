timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alternating or errors or c1_batch" 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ktiming on ', d['value'], d['ms_per_step'], d['encode_gbs'], d['decode_gbs'])"
DRICE_BENCH_NO_KTIMING=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ktiming off', d['value'], d['ms_per_step'], d['encode_gbs'], d['decode_gbs'])"
done
