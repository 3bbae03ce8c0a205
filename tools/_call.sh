mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/s3_gputests.log 2>&1
tail -3 gpurun_out/s3_gputests.log
( time timeout 400 python bench.py ) > gpurun_out/s3_bench_n1.json 2> gpurun_out/s3_bench_n1.err
tail -c 600 gpurun_out/s3_bench_n1.json; tail -3 gpurun_out/s3_bench_n1.err
( time timeout 200 python tools/profile_all.py gpu polyprism_gz ) > gpurun_out/s3_prof.log 2>&1
tail -4 gpurun_out/s3_prof.log
