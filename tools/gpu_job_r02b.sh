set -x
python -m pytest tests/test_gpu_host.py -m gpu -q > gpurun_out/r02_gputests_host_shared.log 2>&1; tail -8 gpurun_out/r02_gputests_host_shared.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke_v4.log 2>&1; tail -12 gpurun_out/r02_smoke_v4.log
timeout 900 python bench.py > gpurun_out/r02_bench_v4_1gpu.json 2> gpurun_out/r02_bench_v4_1gpu.err; cat gpurun_out/r02_bench_v4_1gpu.json
