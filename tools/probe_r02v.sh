mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== register twiddles (default), run $rep"; timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
echo "== shared-memory twiddle table (CPB_Z_TWREG=0 build), run $rep"; CPB200_LIB=$PWD/cpmd_b200/libcpb200_tw0.so timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
done
} > gpurun_out/r02v_probe_twreg.txt 2>&1
cat gpurun_out/r02v_probe_twreg.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_entry_points_match or large_meshes or production_shape or golden" > gpurun_out/r02v_pytest.log 2>&1; tail -3 gpurun_out/r02v_pytest.log
