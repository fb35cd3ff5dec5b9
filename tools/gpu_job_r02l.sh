set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests_full_suite_v5.log 2>&1; tail -6 gpurun_out/r02_gputests_full_suite_v5.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v7.log 2>&1; tail -4 gpurun_out/r02_smoke_v7.log
