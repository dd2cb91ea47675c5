mkdir -p gpurun_out
{
echo "== baseline (CPB_XW=1)"; CPB_XW=1 timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -1
for k in 1 2 3; do echo "== debug knob $k (1: T1 stores stay in L2, 2: gather from a 16 KB window, 3: both), CPB_XW=1"; CPB_XW=1 CPB200_LIB=$PWD/cpmd_b200/libcpb200_dbg$k.so timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -1; done
} > gpurun_out/r02o_probe_xinv_knobs.txt 2>&1
cat gpurun_out/r02o_probe_xinv_knobs.txt
