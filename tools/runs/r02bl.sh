timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "scattering_one_read or ltimes_abi or moments_tensor_core_shapes or folded" > gpurun_out/r02bl_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02bl_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|memcheck rc" gpurun_out/r02bl_memcheck.log | tail -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "scattering_one_read and (0-DGZ or 2-DGZ or 3-DGZ) or ltimes_abi" > gpurun_out/r02bl_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02bl_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|racecheck rc" gpurun_out/r02bl_racecheck.log | tail -5
