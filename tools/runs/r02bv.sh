timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02bv_bench_config2_DGZ.json 2> gpurun_out/r02bv_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02bv_bench_config2_DGZ.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"], d["particles_last"])
PY
tail -2 gpurun_out/r02bv_bench.err
