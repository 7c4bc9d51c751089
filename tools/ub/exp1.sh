# Session-3 experiments of round 2 (profiles/r02_experiments.txt, last section).  Run under gpurun from the repo root.
# load-only replicas of the tcgen05 kernels' CTA -> address mapping
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ub/tma_map tools/ub/tma_map.cu -lcuda && tools/ub/tma_map
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ub/tma_map2 tools/ub/tma_map2.cu -lcuda && tools/ub/tma_map2
# phase B timeline (bit 10 of SDN_UMMA_DBG_NOSHARED = trace); + 16: epilogue loads/stores off (wrong results)
SDN_UMMA_DBG_NOSHARED=1024 python tools/gpu_accum_trace.py 64 3000
SDN_UMMA_UNTILE=0 SDN_UMMA_DBG_NOSHARED=1024 python tools/gpu_accum_trace.py 64 3000
SDN_UMMA_UNTILE=0 SDN_UMMA_DBG_NOSHARED=1040 python tools/gpu_accum_trace.py 64 3000
# per-kernel times of the chain with and without the tile scratch; balanced phase B; partial sums through the scratch
python tools/gpu_umma_l2keep.py 64 3000
SDN_UMMA_UNTILE=0 python tools/gpu_umma_l2keep.py 64 3000
SDN_UMMA_BALANCED=1 SDN_UMMA_UNTILE=0 python tools/gpu_umma_l2keep.py 64 3000
SDN_UMMA_UNTILE_PARTIAL=1 python tools/gpu_partial_probe.py 64 375
