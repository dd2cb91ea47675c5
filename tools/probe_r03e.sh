mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== k_y_fwd cp.async slots (default), run $rep"; timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
echo "== k_y_fwd register prefetch (CPB_YFWD_ASYNC=0 build), run $rep"; CPB200_LIB=$PWD/cpmd_b200/libcpb200_yf0.so timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
done
echo "== CPB_PDL=0x3f (y_fwd may start early too)"; CPB_PDL=0x3f timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
} > gpurun_out/r03e_probe_yfwd.txt 2>&1
cat gpurun_out/r03e_probe_yfwd.txt
