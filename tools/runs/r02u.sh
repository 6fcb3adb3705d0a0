for il in 0 1; do
  echo "== INTERLEAVE=$il"
  KB200_PENCIL_INTERLEAVE=$il timeout 300 python tools/gpu_probe.py config2:GZD config2:ZGD 2>&1 | grep -E "config2|Sweep"
done
KB200_PENCIL_INTERLEAVE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "sweep or solve or golden or population" 2>&1 | tail -2
