# round 2, visit a: where the thread-per-line sweep's time goes (timing experiments) + ncu evidence for the element-fastest sweep
TAG=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
for x in 0 8 5 2 7 15 31; do
  echo "== ZEXP=$x" >> gpurun_out/${TAG}_zline_exp.log
  KB200_SWEEP_IROW=0 KB200_ZEXP=$x timeout 300 python tools/gpu_probe.py config2:DGZ 2>&1 | grep -E "SweepSolver|config2" >> gpurun_out/${TAG}_zline_exp.log
done
cat gpurun_out/${TAG}_zline_exp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_elem' --launch-skip 15 -c 15 -o gpurun_out/${TAG}_full_config2_GZD_sweep_elem python tools/gpu_probe.py config2:GZD > gpurun_out/${TAG}_full_elem.log 2>&1
tail -3 gpurun_out/${TAG}_full_elem.log
