set -x
python -m pytest tests/test_gpu_cellpop.py tests/test_gpu_host.py -m gpu -q -x -k "time_points or time_course or golden or chunks" > gpurun_out/r02_gputests_per_cell.log 2>&1; tail -25 gpurun_out/r02_gputests_per_cell.log
