mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== prime-factor DFT-12 (default), run $rep"; timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
echo "== Cooley-Tukey DFT-12 (CPB_NO_PFA build), run $rep"; CPB200_LIB=$PWD/cpmd_b200/libcpb200_nopfa.so timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
done
} > gpurun_out/r03d_probe_pfa.txt 2>&1
cat gpurun_out/r03d_probe_pfa.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03d_pytest_gpu.log 2>&1; tail -3 gpurun_out/r03d_pytest_gpu.log
