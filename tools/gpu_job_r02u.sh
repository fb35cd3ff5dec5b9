set -x
python tools/cellpop_env_variants.py --n 12 --cells 10000 "LOCKSTEP_TEAM=6" "LOCKSTEP_TEAM=4" "LOCKSTEP_TEAM=3" "LOCKSTEP_TEAM=2" > gpurun_out/r02_cellpop12_variants_lockstep_team.log 2>&1; cat gpurun_out/r02_cellpop12_variants_lockstep_team.log
