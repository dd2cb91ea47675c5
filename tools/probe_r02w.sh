mkdir -p gpurun_out
{
echo "== baseline"; timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -1
for k in 4 8 16 28; do echo "== k_x_fwd_m knob $k (4: c0/c2 gathers from a 16 KB window, 8: c2 stores into a 16 KB window, 16: T1 reads L2-resident, 28: all)"; CPB200_LIB=$PWD/cpmd_b200/libcpb200_dbg$k.so timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -1; done
} > gpurun_out/r02w_probe_xfwd_knobs.txt 2>&1
cat gpurun_out/r02w_probe_xfwd_knobs.txt
