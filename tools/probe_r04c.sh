# A/B on one box: gathers of the mirror-pair x kernels with an L2 evict-last hint (libcpb200_keep.so) against the
# default build; then the DRAM bytes of the x kernels of both builds (ncu, 3 metrics)
mkdir -p gpurun_out
{
for rep in 1 2; do
  echo "== default rep $rep"; timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
  echo "== evict-last gathers rep $rep"; CPB200_LIB=cpmd_b200/libcpb200_keep.so timeout 300 python tools/gpu_probe.py 192 256 32 3 2>&1 | tail -3
done
for lib in libcpb200.so libcpb200_keep.so; do
  echo "== ncu dram bytes, $lib"
  CPB200_LIB=cpmd_b200/$lib timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --kernel-name regex:"k_x_inv_m|k_x_fwd_m" --launch-skip 6 --launch-count 6 --csv python tools/gpu_probe.py 192 128 32 1 2>&1 | grep -E "k_x_(inv|fwd)_m" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200
done
} > gpurun_out/r04c_probe_gather_keep.txt 2>&1
cat gpurun_out/r04c_probe_gather_keep.txt
