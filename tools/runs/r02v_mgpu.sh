N=8
TAG=${1:-r02v}
mkdir -p gpurun_out
for wl in config4 config5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/${TAG}_mgpu${N}_bench_${wl}.json 2> gpurun_out/${TAG}_mgpu${N}_bench_${wl}.err
  tail -2 gpurun_out/${TAG}_mgpu${N}_bench_${wl}.err
done
python - <<PY
import json
for wl in ("config4","config5"):
    f="gpurun_out/${TAG}_mgpu8_bench_%s.json"%wl
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(wl, d["ms_per_step"], {k:round(v["ms"],3) for k,v in d["per_kernel"].items()}, d["roofline"]["frac"], d["sweep_kernel"], d.get("parity",{}).get("max_rel_err"))
    except Exception as e: print(f, "ERR", e)
PY
