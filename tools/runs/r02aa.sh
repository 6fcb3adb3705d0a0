timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > gpurun_out/r02aa_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02aa_pytest_gpu.log
tail -3 gpurun_out/r02aa_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tools/gpu_probe.py config2:DGZ 2>&1 | grep -E "config|Sweep"
python - <<'PY'
import sys; sys.path.insert(0,'.')
import kripke_b200 as kb, ctypes as C
kb.init_device(0)
A=kb.abi(); A.kb200_last_sweep_kernel.restype=C.c_char_p
for z in ("48,48,48","96,32,32","24,24,24"):
    p=kb.Problem(f"--zones {z} --groups 32 --quad 96 --legendre 4 --niter 2"); parts=p.solve(); print(z, A.kb200_last_sweep_kernel().decode(), parts[-1], "SweepSolver timer", p.timer("SweepSolver")); p.close()
PY
