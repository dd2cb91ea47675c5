mkdir -p gpurun_out
{
for px in 0.25 0.5 1 2 4 8; do echo "== CPB_PROLOGUE_X=$px"; CPB_PROLOGUE_X=$px timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -2; done
for v in _zr1 _yi1; do echo "== lib$v"; CPB200_LIB=$PWD/cpmd_b200/libcpb200$v.so timeout 300 python tools/gpu_probe.py 192 256 32 2 2>&1 | tail -2; done
} > gpurun_out/r02d_probe.txt 2>&1
cat gpurun_out/r02d_probe.txt
