# One GPU-box visit: parity tests, the bench lines, the ncu launch list and the ncu --set full captures.
# The .ncu-rep files are summarised ON the box and removed (gpurun brings back at most 64 MiB).
# usage: gpurun --timeout 2400 -- 'bash tools/run_gpu_round.sh TAG COMMIT'
TAG=${1:-r02x}
COMMIT=${2:-unknown}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_smi.log 2>&1
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
tail -3 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_config2_DGZ.json 2> $O/${TAG}_bench.err
cut -c1-400 $O/${TAG}_bench_config2_DGZ.json
for lay in GZD ZGD; do
  timeout 600 python bench.py --steps 10 --warmup 3 --layout $lay --no-cpu-baseline > $O/${TAG}_bench_config2_${lay}.json 2>> $O/${TAG}_bench.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_reference_arm.json 2>> $O/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_config2_DGZ.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sweep_|moments_|ltimes_slab|scatter_|population' --launch-skip 6 -c 6 -o /tmp/${TAG}_full_config2_DGZ python tools/gpu_probe.py config2:DGZ > $O/${TAG}_full_dgz.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_full_config2_DGZ.ncu-rep > $O/${TAG}_ncu_full_config2_DGZ_summary.txt 2>&1
python tools/ncu_traffic.py /tmp/${TAG}_full_config2_DGZ.ncu-rep config2:DGZ $O/${TAG}_ncu_traffic.json $COMMIT > /dev/null 2>&1
python tools/ncu_opcodes.py /tmp/${TAG}_full_config2_DGZ.ncu-rep 30 > $O/${TAG}_ncu_opcodes_config2_DGZ_first_kernel.txt 2>&1
for lay in GZD ZGD; do
  timeout 600 ncu --set full --clock-control none -k regex:'sweep_pencil' --launch-skip 11 -c 11 -o /tmp/${TAG}_full_config2_${lay} python tools/gpu_probe.py config2:$lay > $O/${TAG}_full_${lay}.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_full_config2_${lay}.ncu-rep > $O/${TAG}_ncu_full_config2_${lay}_sweep_pencil_summary.txt 2>&1
  python tools/ncu_traffic.py /tmp/${TAG}_full_config2_${lay}.ncu-rep config2:$lay $O/${TAG}_ncu_traffic.json $COMMIT > /dev/null 2>&1
done
timeout 600 python tools/gpu_probe.py config2:DGZ config2:GDZ config2:GZD config2:ZGD config2:DZG config2:ZDG config3:DGZ config1:DGZ config4:DGZ config5:DGZ > $O/${TAG}_probe.log 2>&1
grep -E "config|total" $O/${TAG}_probe.log | cut -c1-120
du -sh $O
