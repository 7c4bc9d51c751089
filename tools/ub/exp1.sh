timeout 120 python tools/gpu_umma_l2keep.py 64 3000 2>&1 | tail -1
SDN_UMMA_ZREDUCE=1 timeout 120 python tools/gpu_umma_l2keep.py 64 3000 2>&1 | tail -1
timeout 120 python tools/gpu_umma_l2keep.py 128 30000 2>&1 | tail -1
timeout 120 python tools/gpu_umma_l2keep.py 16 515 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SDN_FLASH_WINDOW=6 SDN_FLASH_NOCOOP=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_flash -s 2 -c 1 --csv --log-file gpurun_out/r02_flash2_cfg3_w6_dram.csv python tools/gpu_flash_probe.py 64 3000 near 3.15 0 > /dev/null 2>&1
SDN_FLASH_WINDOW=7 SDN_FLASH_NOCOOP=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_flash -s 2 -c 1 --csv --log-file gpurun_out/r02_flash2_cfg3_w7_dram.csv python tools/gpu_flash_probe.py 64 3000 near 3.15 0 > /dev/null 2>&1
tail -3 gpurun_out/r02_flash2_cfg3_w6_dram.csv gpurun_out/r02_flash2_cfg3_w7_dram.csv
