mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench_cfg3_final.json 2> gpurun_out/bench.err && echo bench ok
# launch list of the same command under ncu (cold-cache, serialised: shares must agree, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_cfg3_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo ncu-list rc=$?
# full captures: the two-phase chain, the stream kernel, the one-pass kernel (plain launch for ncu)
ncu --set full --clock-control none --import-source on -k regex:k_umma -s 12 -c 4 -o gpurun_out/r02_umma_cfg3 -f python tools/gpu_umma_l2keep.py 64 3000 > gpurun_out/ncu_umma.log 2>&1; echo ncu-umma rc=$?
SDN_FLASH_NOCOOP=1 ncu --set full --clock-control none --import-source on -k regex:k_flash -s 2 -c 1 -o gpurun_out/r02_flash2_cfg3 -f python tools/gpu_flash_probe.py 64 3000 near 3.15 0 > gpurun_out/ncu_flash.log 2>&1; echo ncu-flash rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 4 -c 2 -o gpurun_out/r02_stream_q1 -f python tools/gpu_stream_probe.py > gpurun_out/ncu_stream.log 2>&1; echo ncu-stream rc=$?
ls -la gpurun_out/*.ncu-rep
