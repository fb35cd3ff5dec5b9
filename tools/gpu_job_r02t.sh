set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v9.log 2>&1; tail -2 gpurun_out/r02_smoke_v9.log | cut -c1-150
python bench.py > gpurun_out/r02_bench_v7_1gpu.json 2> gpurun_out/r02_bench_v7_1gpu.err; tail -c 200 gpurun_out/r02_bench_v7_1gpu.json; tail -2 gpurun_out/r02_bench_v7_1gpu.err
