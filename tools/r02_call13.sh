set -x
for v in base stag900 stag2800; do
  GPRB_LIB=variants/libgpr_b200_$v.so timeout 200 python tools/perf_s5.py 100 1 3 >> gpurun_out/r02_c13_perf.log 2>&1
done
cat gpurun_out/r02_c13_perf.log | cut -c1-260
