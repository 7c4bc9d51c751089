timeout 100 python tools/gpu_partial_probe.py 64 375 2>&1 | tail -1
SDN_UMMA_UNTILE=0 timeout 100 python tools/gpu_partial_probe.py 64 375 2>&1 | tail -1
timeout 100 python tools/gpu_partial_probe.py 64 3000 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "umma or fused or golden or sparse or spell" 2>&1 | tail -3
