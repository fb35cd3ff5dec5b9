set -x
python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_cellpop.py -m gpu -q -x -k "multidevice or device or chunks or rank or shard" > gpurun_out/r02_gputests_2gpu.log 2>&1; tail -8 gpurun_out/r02_gputests_2gpu.log
