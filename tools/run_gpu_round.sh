# One GPU-box visit: parity tests, the bench line, the ncu launch list and one ncu --set full capture.
# usage: gpurun --timeout 1500 -- 'bash tools/run_gpu_round.sh TAG'
TAG=${1:-r01x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; echo "pytest rc=$?") > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sweep_|moments_|scatter_|population|transpose' --launch-skip 8 -c 8 -o gpurun_out/${TAG}_full_config2_DGZ python tools/gpu_probe.py config2:DGZ > gpurun_out/${TAG}_full.log 2>&1
tail -3 gpurun_out/${TAG}_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config2_DGZ.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config2_DGZ.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_config2_DGZ.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_reference_arm.json
timeout 600 python tools/gpu_probe.py config2:GZD config2:ZGD config3:DGZ config1:DGZ config5:DGZ > gpurun_out/${TAG}_probe.log 2>&1
cat gpurun_out/${TAG}_probe.log
