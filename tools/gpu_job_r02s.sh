set -x
python -m pytest tests/test_gpu_cellpop.py tests/test_gpu_host.py -m gpu -q -x -k "mitotic or dividing or golden" > gpurun_out/r02_gputests_mitotic.log 2>&1; tail -25 gpurun_out/r02_gputests_mitotic.log | cut -c1-300
