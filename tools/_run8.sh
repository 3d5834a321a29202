N=$1
PORT=29611
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "${@:2}"; PORT=$((PORT+1)); }
run 900 --workload c5 --steps 5 --warmup 3 > gpurun_out/r2h_c5_${N}gpu.json 2> gpurun_out/r2h_c5_${N}gpu.err; echo "c5 N=$N rc=$?"
run 900 --steps 10 --warmup 3 > gpurun_out/r2h_c2_${N}gpu.json 2> gpurun_out/r2h_c2_${N}gpu.err; echo "c2 N=$N rc=$?"
