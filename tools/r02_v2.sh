set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_v2_tests.log 2>&1; echo rc=$? >> gpurun_out/r02_v2_tests.log
timeout 280 ncu --set full --clock-control none --profile-from-start off -k regex:prep_rows -c 2 -o /tmp/r02_pack2 -f python tools/profile_misc.py pack > gpurun_out/r02_v2_ncu_pack.log 2>&1
python tools/ncu_kernels_table.py /tmp/r02_pack2.ncu-rep > gpurun_out/r02_misc_pack_v2.txt 2>&1; rm -f /tmp/r02_pack2.ncu-rep
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_v2_smoke.log 2>&1; echo rc=$? >> gpurun_out/r02_v2_smoke.log
tail -n 3 gpurun_out/r02_v2_tests.log; cat gpurun_out/r02_misc_pack_v2.txt | cut -c1-200; tail -n 2 gpurun_out/r02_v2_smoke.log
