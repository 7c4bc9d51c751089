mkdir -p gpurun_out
timeout 120 python tools/gpu_umma_l2keep.py 64 3000 2>&1 | tail -1
timeout 120 python tools/gpu_umma_l2keep.py 64 375 2>&1 | tail -1
timeout 120 python tools/gpu_umma_l2keep.py 128 30000 2>&1 | tail -1
timeout 120 python tools/gpu_flash_probe.py 64 3000 near 3.15 1 2>&1 | tail -1 | cut -c 60-
