set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_full_suite_v6.log 2>&1; tail -6 gpurun_out/r02_gputests_full_suite_v6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v8.log 2>&1; tail -3 gpurun_out/r02_smoke_v8.log
python bench.py > gpurun_out/r02_bench_v6_1gpu.json 2> gpurun_out/r02_bench_v6_1gpu.err; tail -c 300 gpurun_out/r02_bench_v6_1gpu.json; tail -2 gpurun_out/r02_bench_v6_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_v6.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_v6_ncu.log 2>&1; tail -2 gpurun_out/r02_bench_v6_ncu.log | cut -c1-200
