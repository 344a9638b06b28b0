set -x
B="python bench.py --steps 1 --warmup 3 --pairs 2048 --no-cpu-baseline --eager-pairs 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/bench_under_ncu_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffn_pair -s 36 -c 2 -o gpurun_out/prof_ffn_r2 -f $B > gpurun_out/ncu_ffn_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_tma_kernel -s 60 -c 8 -o gpurun_out/prof_ltma_r2 -f $B > gpurun_out/ncu_ltma_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:recon_pool_tiled -s 8 -c 1 -o gpurun_out/prof_recon_r2 -f $B > gpurun_out/ncu_recon_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_mma_kernel -s 6 -c 1 -o gpurun_out/prof_attn_self_r2 -f $B > gpurun_out/ncu_attn_self_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_mma_heads -s 10 -c 1 -o gpurun_out/prof_attn_t2v_r2 -f $B > gpurun_out/ncu_attn_t2v_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dec_cross_mma -s 2 -c 1 -o gpurun_out/prof_dec_cross_r2 -f $B > gpurun_out/ncu_dec_cross_r2.log 2>&1
wc -l gpurun_out/launches_r2.csv
ls -la gpurun_out/*_r2.ncu-rep
