# ncu --set full of the non-K_ff kernels, one phase per run, only the library's own kernels of the captured pass
# (cudaProfilerStart/Stop around the second pass of tools/profile_misc.py).  The reports are summarised ON the GPU box
# (tools/ncu_kernels_table.py) and deleted: gpurun copies back at most 64 MiB.
set -x
run() {  # phase, kernel regex, launch count
  timeout 280 ncu --set full --clock-control none --profile-from-start off -k "$2" -c $3 -o /tmp/r02_misc_$1 -f \
      python tools/profile_misc.py $1 > gpurun_out/r02_ncu_$1.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_ncu_$1.log
  python tools/ncu_kernels_table.py /tmp/r02_misc_$1.ncu-rep > gpurun_out/r02_misc_$1.txt 2>&1
  rm -f /tmp/r02_misc_$1.ncu-rep
}
run so3 'regex:so3_' 8
run pack 'regex:prep_rows' 2
run kef 'regex:cov_mma' 2
run kee 'regex:cov_mma' 2
run lml 'regex:trace_kernel|mirror_upper|lml_terms|final_sum|cov_mma' 12
run predict 'regex:predict_rows|cov_mma|kee_diag' 16
cat gpurun_out/r02_misc_*.txt | cut -c1-220
du -sh gpurun_out
