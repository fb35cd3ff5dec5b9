mkdir -p gpurun_out; rm -f gpurun_out/poppk_occ_ref.npy
L=gpurun_out/r02_poppk_occupancy_variants.log; : > $L
python tools/poppk_occupancy_variants.py 0,384 2>&1 | tee -a $L
BCM3B200_LIB=$PWD/bcm3_b200/libbcm3b200_r152s416.so python tools/poppk_occupancy_variants.py 416,128 2>&1 | tee -a $L
BCM3B200_LIB=$PWD/bcm3_b200/libbcm3b200_r144s448.so python tools/poppk_occupancy_variants.py 448,64 2>&1 | tee -a $L
BCM3B200_LIB=$PWD/bcm3_b200/libbcm3b200_r144s224.so python tools/poppk_occupancy_variants.py 224 2>&1 | tee -a $L
BCM3B200_LIB=$PWD/bcm3_b200/libbcm3b200_r160s384.so python tools/poppk_occupancy_variants.py 0,384 2>&1 | tee -a $L
