# round-2 final validation: GPU tests, bench lines (cfg3 default, cfg1), launch list under ncu
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 200 python bench.py --workload cfg1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_cfg1.json 2>> gpurun_out/bench.err
python - <<'PY'
import json
for f in ("r02_bench_cfg3", "r02_bench_cfg1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"]["frac"], json.dumps(d["e2e"])[:400], d.get("parity_check"), d["clocks"])
        print({k: v["avg_ms"] for k, v in d["roofline"]["kernels"].items()})
    except Exception as ex:
        print(f, "failed", ex)
PY
