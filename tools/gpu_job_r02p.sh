set -x
python -m pytest tests/test_gpu_host.py -m gpu -q -x > gpurun_out/r02_gputests_host_final.log 2>&1; tail -4 gpurun_out/r02_gputests_host_final.log
