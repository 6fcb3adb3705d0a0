timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02ca_mgpu2_bench.json 2> gpurun_out/r02ca_mgpu2_bench.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02ca_mgpu2_bench.json').read().strip().splitlines()[-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("parity",{}).get("max_rel_err"), d.get("scattering_kernel"))
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/r02ca_mgpu2_bench.err | cut -c1-200
