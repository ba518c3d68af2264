set -x
for v in base staged; do
  GPRB_LIB=variants/libgpr_b200_$v.so ncu --set full --import-source on --clock-control none -k regex:cov_mma -s 1 -c 1 -o gpurun_out/r02_kff_$v -f python tools/perf_s5.py 100 1 1 > gpurun_out/r02_c3_ncu_$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
