for v in 0 1 0 1; do echo interleave=$v; SDN_UMMA_INTERLEAVE_K=$v timeout 120 python tools/gpu_umma_l2keep.py 64 3000 2>&1 | tail -1 | cut -c 50-; done
SDN_UMMA_INTERLEAVE_K=1 timeout 120 python tools/gpu_umma_l2keep.py 128 30000 2>&1 | tail -1 | cut -c 50-
SDN_UMMA_INTERLEAVE_K=1 timeout 120 python tools/gpu_umma_l2keep.py 64 375 2>&1 | tail -1 | cut -c 50-
