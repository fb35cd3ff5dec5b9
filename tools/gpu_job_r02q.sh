set -x
python tools/poppk_tail_blocks.py 12500 25000 100000 > gpurun_out/r02_poppk_tail_blocks.log 2>&1; cat gpurun_out/r02_poppk_tail_blocks.log
