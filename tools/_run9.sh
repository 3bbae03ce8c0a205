N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r02_n$N.err | grep '^{' > gpurun_out/r02_bench_n$N.json
wc -c gpurun_out/r02_bench_n$N.json
