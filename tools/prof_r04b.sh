# end-of-round ncu --set full capture of every hot-path kernel (one serialised iteration after a warm one):
# DRAM bytes per launch for profiles/r04b_traffic.json, pipe utilisation and stall hot spots
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none \
  --kernel-name regex:"k_x_inv_m|k_y_inv|k_z_rho|k_z_vpsi|k_y_fwd|k_x_fwd_m" --launch-skip 16 --launch-count 16 \
  -o gpurun_out/prof_r04b_all -f python tools/gpu_probe.py 192 128 32 1 > gpurun_out/prof_r04b_all.log 2>&1
tail -3 gpurun_out/prof_r04b_all.log
ls -la gpurun_out/prof_r04b_all.ncu-rep
