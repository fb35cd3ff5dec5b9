set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -q -k "single_patient or batched_equals_serial or shared_integration" > gpurun_out/r02_gputests_single_patient.log 2>&1; tail -8 gpurun_out/r02_gputests_single_patient.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_full_suite_v2.log 2>&1; tail -6 gpurun_out/r02_gputests_full_suite_v2.log
