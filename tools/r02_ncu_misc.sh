# ncu --set full of the non-K_ff kernels, one phase per run, only the library's own kernels of the captured pass.
# Reports are read back with tools/ncu_kernels_table.py; keep gpurun_out/ under 64 MiB (no --import-source here).
set -x
K_so3='regex:so3_'
K_pack='regex:prep_rows'
K_kef='regex:cov_mma'
K_kee='regex:cov_mma'
K_lml='regex:trace_kernel|mirror_upper|lml_terms|final_sum|cov_mma'
K_predict='regex:predict_rows|cov_mma|kee_diag'
for ph in so3 pack kef kee lml predict; do
  eval k=\$K_$ph
  timeout 300 ncu --set full --clock-control none --profile-from-start off -k "$k" -c 48 -o gpurun_out/r02_misc_$ph -f \
      python tools/profile_misc.py $ph > gpurun_out/r02_ncu_$ph.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_ncu_$ph.log
done
du -sh gpurun_out; ls -la gpurun_out
