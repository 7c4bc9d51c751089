for v in 1 0; do echo merged=$v; SDN_UMMA_MERGED_TMA=$v timeout 120 python tools/gpu_umma_l2keep.py 64 3000 2>&1 | tail -1 | cut -c 50-; done
timeout 120 python tools/gpu_umma_l2keep.py 128 30000 2>&1 | tail -1 | cut -c 50-
timeout 120 python tools/gpu_umma_l2keep.py 64 375 2>&1 | tail -1 | cut -c 50-
timeout 120 python tools/gpu_umma_l2keep.py 16 515 2>&1 | tail -1 | cut -c 50-
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
