for a in 1 0; do for i in 0 1; do
  echo "ALIGN=$a INTERLEAVE=$i"
  KB200_PENCIL_ALIGN=$a KB200_PENCIL_INTERLEAVE=$i timeout 60 python tools/gpu_probe.py config2:GZD 2>&1 | grep -E "Sweep"
done; done | tee gpurun_out/r02aq_pencil_gzd.log
