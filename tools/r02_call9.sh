set -x
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r02_c9_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_c9_tests.log
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:so3_ -c 8 -o /tmp/r02_so3v2 -f python tools/profile_misc.py so3 > gpurun_out/r02_c9_ncu_so3.log 2>&1
python tools/ncu_kernels_table.py /tmp/r02_so3v2.ncu-rep > gpurun_out/r02_misc_so3_v2.txt 2>&1; rm -f /tmp/r02_so3v2.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cov_mma -s 1 -c 1 -o /tmp/r02_kff_s5 -f python tools/perf_s5.py 340 1 1 > gpurun_out/r02_c9_ncu_kff.log 2>&1
python tools/ncu_summary.py /tmp/r02_kff_s5.ncu-rep gpurun_out/r02_kff_s5_ncu_summary.txt > /dev/null 2>&1; rm -f /tmp/r02_kff_s5.ncu-rep
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_c9_bench.log 2> gpurun_out/r02_c9_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_s5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-predict --no-e2e --no-small --no-s4 > gpurun_out/r02_c9_ncu_launches.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r02_c9_tests.log | tail -n 20; cat gpurun_out/r02_misc_so3_v2.txt | cut -c1-200; tail -n 4 gpurun_out/r02_c9_bench.err; du -sh gpurun_out
