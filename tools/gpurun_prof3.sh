set -x
B="python bench.py --steps 1 --warmup 3 --pairs 1024 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:attn_tcp -s 12 -c 1 -o gpurun_out/prof_attnp_r1f -f $B > gpurun_out/ncu_attnp_f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 3 --pairs 2048 --no-cpu-baseline > gpurun_out/bench_under_ncu_f.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python sweep.py --pairs 65536 2>&1 | tail -2
