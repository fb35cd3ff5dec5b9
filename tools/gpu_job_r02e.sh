set -x
python tools/cellpop_env_variants.py --n 50 --cells 6000 --decades 4.0 "SOLVE_SKIP=0" "GROUP_WARPS=8" "GROUP_WARPS=3,GROUP_MIN_BLOCKS=3,GROUP_LOCKSTEP=1" "GROUP_WARPS=3,GROUP_MIN_BLOCKS=3" "GROUP_WARPS=4,GROUP_MIN_BLOCKS=2,GROUP_LOCKSTEP=1" "GROUP_WARPS=4,GROUP_MIN_BLOCKS=2" > gpurun_out/r02_cellpop50_variants_solve_skip.log 2>&1
cat gpurun_out/r02_cellpop50_variants_solve_skip.log
python tools/cellpop_env_variants.py --n 24 --cells 6000 --decades 4.0 "SOLVE_SKIP=0" > gpurun_out/r02_cellpop24_variants_solve_skip.log 2>&1
cat gpurun_out/r02_cellpop24_variants_solve_skip.log
python -m pytest tests/test_gpu_cellpop.py -m gpu -q -x > gpurun_out/r02_gputests_cellpop_solve_skip.log 2>&1; tail -5 gpurun_out/r02_gputests_cellpop_solve_skip.log
