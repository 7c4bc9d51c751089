timeout 120 python tools/gpu_umma_l2keep.py 64 3000 2>&1 | tail -1 | cut -c 50-
SDN_UMMA_UNTILE=0 timeout 120 python tools/gpu_umma_l2keep.py 64 3000 2>&1 | tail -1 | cut -c 50-
timeout 120 python tools/gpu_umma_l2keep.py 128 3000 2>&1 | tail -1 | cut -c 50-
timeout 120 python tools/gpu_umma_l2keep.py 64 375 2>&1 | tail -1 | cut -c 50-
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "umma or fused or golden or host" 2>&1 | tail -3
