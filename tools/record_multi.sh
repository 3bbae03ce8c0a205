# Records of one N-GPU box (gpurun --gpus N -- bash tools/record_multi.sh N): default bench, the north-star
# runs, the full C4 matrix, row-sharded normal equations, NCCL check -> gpurun_out/r02_*
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r02_n$N.err | grep '^{' > gpurun_out/r02_bench_n$N.json
$TR --master-port 29513 bench.py --gpus $N --workload c5tfull --steps 1 --warmup 3 --no-cpu --no-e2e --no-configs 2>>gpurun_out/r02_n$N.err | grep '^{' > gpurun_out/r02_bench_n${N}_c5tfull.json
$TR --master-port 29514 bench.py --gpus $N --workload c5full --steps 1 --warmup 3 --no-cpu --no-e2e --no-configs 2>>gpurun_out/r02_n$N.err | grep '^{' > gpurun_out/r02_bench_n${N}_c5full.json
$TR --master-port 29515 bench.py --gpus $N --workload c4full --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs 2>>gpurun_out/r02_n$N.err | grep '^{' > gpurun_out/r02_bench_n${N}_c4full.json
$TR --master-port 29516 tools/bench_normal.py 2>>gpurun_out/r02_n$N.err | grep '^{' > gpurun_out/r02_normal_n$N.json
$TR --master-port 29517 tests/dist_nccl_check.py 2>&1 | grep "dist_nccl_check" > gpurun_out/r02_nccl_check_n$N.txt
tail -3 gpurun_out/r02_n$N.err; wc -c gpurun_out/r02_*n$N*; cat gpurun_out/r02_nccl_check_n$N.txt
