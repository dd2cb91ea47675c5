mkdir -p gpurun_out
{
echo "== 320^3 x 32 states, batch 16"; timeout 300 python tools/gpu_probe.py 320 32 16 2 2>&1 | tail -3
echo "== 288^3 x 32 states, batch 16"; timeout 300 python tools/gpu_probe.py 288 32 16 2 2>&1 | tail -2
echo "== 256^3 x 64 states, batch 16"; timeout 300 python tools/gpu_probe.py 256 64 16 2 2>&1 | tail -2
} > gpurun_out/r02u_probe_large.txt 2>&1
cat gpurun_out/r02u_probe_large.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_meshes or largest_lengths" > gpurun_out/r02u_pytest_large.log 2>&1; tail -3 gpurun_out/r02u_pytest_large.log
