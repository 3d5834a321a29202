N=$1
PORT=29611
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "${@:2}"; PORT=$((PORT+1)); }
run 900 --workload c5 --steps 5 --warmup 3 > gpurun_out/r2_c5_${N}gpu.json 2> gpurun_out/r2_c5_${N}gpu.err; echo "c5 N=$N rc=$?"
run 600 --impl reference --workload c5 --steps 3 --warmup 1 > gpurun_out/r2_c5_${N}gpu_ref.json 2> gpurun_out/r2_c5_${N}gpu_ref.err; echo "c5 ref N=$N rc=$?"
run 900 --steps 10 --warmup 3 > gpurun_out/r2_c2_${N}gpu.json 2> gpurun_out/r2_c2_${N}gpu.err; echo "c2 N=$N rc=$?"
run 600 --impl reference --steps 3 --warmup 1 > gpurun_out/r2_c2_${N}gpu_ref.json 2> gpurun_out/r2_c2_${N}gpu_ref.err; echo "c2 ref N=$N rc=$?"
