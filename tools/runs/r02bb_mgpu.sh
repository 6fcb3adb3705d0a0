# multi-GPU visit after the scattering rewrite: parity against the reference goldens, then the weak-scaling bench line (config 2)
N=${1:-2}
TAG=${2:-r02bb}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/mgpu_check.py > gpurun_out/${TAG}_mgpu${N}_parity.log 2>&1
grep -E "PASS|FAIL|rror" gpurun_out/${TAG}_mgpu${N}_parity.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_mgpu${N}_bench_config2.json 2> gpurun_out/${TAG}_mgpu${N}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_mgpu${N}_bench_config2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], {k:round(v["ms"],3) for k,v in d["per_kernel"].items()}, d["e2e"]["ms_per_step"], d.get("parity",{}).get("max_rel_err"), d.get("scattering_kernel"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/${TAG}_mgpu${N}_bench.err
if [ "$N" = "8" ]; then
for wl in config4 config5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/${TAG}_mgpu${N}_bench_${wl}.json 2>> gpurun_out/${TAG}_mgpu${N}_bench.err
  python - <<PY
import json
f="gpurun_out/${TAG}_mgpu${N}_bench_${wl}.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], {k:round(v["ms"],3) for k,v in d["per_kernel"].items()}, d["roofline"]["frac"])
except Exception as e: print(f, "ERR", e)
PY
done
fi
