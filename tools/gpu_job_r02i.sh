set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_full_suite_v4.log 2>&1; tail -6 gpurun_out/r02_gputests_full_suite_v4.log
python bench.py > gpurun_out/r02_bench_v5_1gpu.json 2> gpurun_out/r02_bench_v5_1gpu.err; tail -c 600 gpurun_out/r02_bench_v5_1gpu.json; tail -3 gpurun_out/r02_bench_v5_1gpu.err
python bench.py --impl reference > gpurun_out/r02_bench_v5_reference.json 2> gpurun_out/r02_bench_v5_reference.err; tail -c 800 gpurun_out/r02_bench_v5_reference.json
