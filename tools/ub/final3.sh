mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/bench.err && echo bench ok
python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_cfg1.json 2>> gpurun_out/bench.err && echo cfg1 ok
ncu --set full --clock-control none --import-source on -k regex:k_umma_dots -s 1 -c 1 -o gpurun_out/r02_umma_dots_cfg3 -f python tools/gpu_umma_l2keep.py 64 3000 > gpurun_out/ncu_dots.log 2>&1; echo ncu-dots rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_cfg3_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo ncu-list rc=$?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
