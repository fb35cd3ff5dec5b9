set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_full_suite_v7.log 2>&1; tail -5 gpurun_out/r02_gputests_full_suite_v7.log
