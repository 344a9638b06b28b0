set -x
B="python bench.py --steps 1 --warmup 3 --pairs 1024 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:recon_pool -s 12 -c 1 -o gpurun_out/prof_recon_r1d -f $B > gpurun_out/ncu_recon_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mha_rows -s 24 -c 1 -o gpurun_out/prof_mharows_r1e -f $B > gpurun_out/ncu_mharows_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_cross -s 6 -c 1 -o gpurun_out/prof_deccross_r1e -f $B > gpurun_out/ncu_deccross_e.log 2>&1
