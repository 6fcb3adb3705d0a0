timeout 300 python -X faulthandler -m pytest tests -m gpu -x -q -k "population or golden or non_uniform or solver" > gpurun_out/r02bw_pytest.log 2>&1; tail -3 gpurun_out/r02bw_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02bw_bench_config2_DGZ.json 2> gpurun_out/r02bw_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02bw_bench_config2_DGZ.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"], d["sweep_kernel"])
PY
tail -2 gpurun_out/r02bw_bench.err
