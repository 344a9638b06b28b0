set -x
B="python bench.py --steps 1 --warmup 3 --pairs 1024 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 3 --pairs 2048 --no-cpu-baseline > gpurun_out/bench_under_ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffn_pair -s 36 -c 2 -o gpurun_out/prof_ffn_r1d -f $B > gpurun_out/ncu_ffn_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 12 -c 1 -o gpurun_out/prof_attn_r1d -f $B > gpurun_out/ncu_attn_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mha_rows -s 24 -c 1 -o gpurun_out/prof_mharows_r1d -f $B > gpurun_out/ncu_mharows_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_tc_kernel -s 300 -c 6 -o gpurun_out/prof_lin_r1d -f $B > gpurun_out/ncu_lin_d.log 2>&1
tail -2 gpurun_out/bench_under_ncu_d.log | cut -c1-300
wc -l gpurun_out/launches_r1d.csv
