timeout 300 python tools/gpu_probe.py config4:DGZ config2:DGZ config1:DGZ config5:DGZ 2>&1 | grep -E "config|Sweep"
timeout 300 python -m pytest tests -m gpu -x -q -k "irow or zone_fastest or golden or solve" 2>&1 | tail -2
