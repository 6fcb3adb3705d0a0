TAG=${1:-r02m}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest rc=$?") > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
python - <<'PY'
import time, sys
sys.path.insert(0, '.')
import kripke_b200 as kb
kb.init_device(0)
import os
for hostgen in ("0", "1"):
    os.environ["KB200_HOST_GENERATOR"] = hostgen
    t0 = time.time(); p = kb.Problem("--zones 128,128,128 --groups 8 --quad 8 --legendre 0 --zset 2,2,2 --gset 1 --dset 8"); t1 = time.time()
    print("generate 128^3 zones, host_generator=%s: %.3f s (Generate timer %.3f s)" % (hostgen, t1 - t0, p.timer("Generate")), flush=True)
    p.close()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_pencil' --launch-skip 15 -c 2 -o gpurun_out/${TAG}_full_config2_GZD_sweep_pencil python tools/gpu_probe.py config2:GZD > gpurun_out/${TAG}_full_gzd.log 2>&1
tail -2 gpurun_out/${TAG}_full_gzd.log
