BCM3B200_LIB=$PWD/bcm3_b200/libbcm3b200_ls1r168.so python tools/variant_bench.py 128,192,384 2>&1 | tee gpurun_out/variants_v4.log
BCM3B200_LIB=$PWD/bcm3_b200/libbcm3b200_ls1r128.so python tools/variant_bench.py 128,256,512 2>&1 | tee -a gpurun_out/variants_v4.log
