set -x
timeout 400 python -m pytest tests/test_gpu_so3.py tests/test_stress.py tests/test_gpu_gp.py -m gpu -q > gpurun_out/r02_c11_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c11_tests.log
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:so3_ -c 8 -o /tmp/r02_so3v3 -f python tools/profile_misc.py so3 > gpurun_out/r02_c11_ncu_so3.log 2>&1
python tools/ncu_kernels_table.py /tmp/r02_so3v3.ncu-rep > gpurun_out/r02_misc_so3_v3.txt 2>&1; rm -f /tmp/r02_so3v3.ncu-rep
tail -n 3 gpurun_out/r02_c11_tests.log; cat gpurun_out/r02_misc_so3_v3.txt | cut -c1-200
