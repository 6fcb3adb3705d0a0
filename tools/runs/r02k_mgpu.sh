# multi-GPU visit: parity of the peer-memory exchange against the reference goldens, then the weak-scaling bench line
N=${1:-2}
TAG=${2:-r02k}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/${TAG}_mgpu${N}_parity.log 2>&1
tail -8 gpurun_out/${TAG}_mgpu${N}_parity.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_mgpu${N}_bench.json 2> gpurun_out/${TAG}_mgpu${N}_bench.err
cat gpurun_out/${TAG}_mgpu${N}_bench.json | cut -c1-3000; tail -5 gpurun_out/${TAG}_mgpu${N}_bench.err
KB200_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_mgpu${N}_bench_nccl.json 2>> gpurun_out/${TAG}_mgpu${N}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_mgpu${N}_bench.json","gpurun_out/${TAG}_mgpu${N}_bench_nccl.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["per_kernel"]["SweepSolver"]["ms"], d["e2e"]["ms_per_step"], d.get("parity",{}).get("max_rel_err"))
    except Exception as e: print(f, "ERR", e)
PY
