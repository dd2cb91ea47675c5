mkdir -p gpurun_out
{
for zw in 1 0; do echo "== CPB_ZW=$zw"; CPB_ZW=$zw timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -3; done
echo "== CPB_ZW=1 96^3"; timeout 300 python tools/gpu_probe.py 96 128 32 2 2>&1 | tail -2
echo "== CPB_ZW=0 96^3"; CPB_ZW=0 timeout 300 python tools/gpu_probe.py 96 128 32 2 2>&1 | tail -2
echo "== CPB_ZW=1 256^3"; timeout 300 python tools/gpu_probe.py 256 64 16 2 2>&1 | tail -2
echo "== CPB_ZW=0 256^3"; CPB_ZW=0 timeout 300 python tools/gpu_probe.py 256 64 16 2 2>&1 | tail -2
} > gpurun_out/r02j_probe_zw.txt 2>&1
cat gpurun_out/r02j_probe_zw.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02j_pytest_gpu.log
