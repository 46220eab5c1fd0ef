./build/tmp/lds128_merge | tee gpurun_out/r02_lds128_merge.txt
