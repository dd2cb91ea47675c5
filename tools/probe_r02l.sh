mkdir -p gpurun_out
{
for xw in 0 2 3; do echo "== CPB_XW=$xw"; CPB_XW=$xw timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -3; done
for px in 0.25 1 4; do echo "== CPB_XW=3 CPB_PROLOGUE_X=$px"; CPB_PROLOGUE_X=$px CPB_XW=3 timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -1; done
} > gpurun_out/r02l_probe_xw.txt 2>&1
cat gpurun_out/r02l_probe_xw.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warp_x" > gpurun_out/r02l_pytest_warp.log 2>&1; tail -3 gpurun_out/r02l_pytest_warp.log
