timeout 120 tools/ub/tma_map 2>&1 | tee gpurun_out/tma_map.txt
