set -x
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
MESM_PROFILE_REPORT=1 python bench.py --no-cpu-baseline --steps 2 2>&1 | grep " ms  n=" > gpurun_out/prof_r1_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 1 --warmup 3 --pairs 2048 --no-cpu-baseline > gpurun_out/bench_under_ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tcp -s 12 -c 1 -o gpurun_out/prof_attnp_r1g -f python bench.py --steps 1 --warmup 3 --pairs 1024 --no-cpu-baseline > gpurun_out/ncu_attnp_g.log 2>&1
