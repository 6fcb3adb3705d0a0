N=${1:-2}
TAG=${2:-r02x}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/mgpu_check.py > gpurun_out/${TAG}_mgpu${N}_parity.log 2>&1
grep -E "PASS|FAIL|rror" gpurun_out/${TAG}_mgpu${N}_parity.log | cut -c1-200
for wl in config5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/${TAG}_mgpu${N}_bench_${wl}.json 2> gpurun_out/${TAG}_mgpu${N}_bench_${wl}.err
  KB200_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/${TAG}_mgpu${N}_bench_${wl}_nccl.json 2>> gpurun_out/${TAG}_mgpu${N}_bench_${wl}.err
done
python - <<PY
import json
for f in ("gpurun_out/${TAG}_mgpu${N}_bench_config5.json","gpurun_out/${TAG}_mgpu${N}_bench_config5_nccl.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], {k:round(v["ms"],3) for k,v in d["per_kernel"].items()}, d.get("parity",{}).get("max_rel_err"))
    except Exception as e: print(f, "ERR", e)
PY
