# Records of one 1-GPU box (gpurun -- bash tools/record_n1.sh): default bench, reference arm, per-cell runs,
# launch list, normal equations, fallback check, harvester, siblings, smoke -> gpurun_out/r02_*
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench_n1.err
for w in c2 c3 c5 c5t; do python bench.py --workload $w --per-cell --steps 3 --no-cpu --no-configs 2>/dev/null | tail -1 >> gpurun_out/r02_bench_percell.jsonl; done
python bench.py --workload c2a --steps 5 --no-cpu --no-configs 2>/dev/null | tail -1 >> gpurun_out/r02_bench_percell.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r02_ncu_launches.log 2>&1
python tools/bench_normal.py > gpurun_out/r02_normal_n1.json 2>/dev/null
python tools/fallback_check.py > gpurun_out/r02_fallback_check.txt 2>&1
python tools/bench_harvester.py > gpurun_out/r02_harvester.json 2>/dev/null
python tools/bench_siblings.py > gpurun_out/r02_siblings.txt 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1
tail -2 gpurun_out/r02_smoke.txt; wc -c gpurun_out/r02_*
