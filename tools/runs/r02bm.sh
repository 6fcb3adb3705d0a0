timeout 600 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 12 python -m pytest tests -m gpu -x -q -k "scattering_one_read and 0-DGZ" > gpurun_out/r02bm_racecheck.log 2>&1
grep -E "=========" gpurun_out/r02bm_racecheck.log | head -60
