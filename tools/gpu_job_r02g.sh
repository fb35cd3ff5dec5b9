set -x
python tools/time_course_timing.py 128 512 2048 > gpurun_out/r02_time_course_timing.log 2>&1; cat gpurun_out/r02_time_course_timing.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v6.log 2>&1; tail -9 gpurun_out/r02_smoke_v6.log
