mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/bench.err && echo cfg3 ok
for w in cfg1 cfg2 cfg4 cfg5; do python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2>> gpurun_out/bench.err && echo $w ok; done
python bench.py --workload cfg3 --sparse --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_cfg3_sparse_accumulate.json 2>> gpurun_out/bench.err && echo sparse ok
python bench.py --workload cfg2-loop --no-cpu-baseline > gpurun_out/r02_bench_cfg2_loop.json 2>> gpurun_out/bench.err && echo loop ok
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_cfg3_reference_arm.json 2>> gpurun_out/bench.err && echo ref ok
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
