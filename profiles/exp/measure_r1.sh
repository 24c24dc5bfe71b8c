# Round-1 measurement pass on one B200 (each ncu run after the same command exited 0 without ncu)
set -x
python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1g_ref.json 2> gpurun_out/bench_r1g_ref.err || exit 1
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/plain_r1g.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/ncu_lg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rmsd_quad -s 13 -c 1 -f -o gpurun_out/prof_r1g python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/ncu_fg.log 2>&1
python profiles/exp/pairs_time.py > gpurun_out/pairs_r1g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pairs_reduce_fast -s 3 -c 1 -f -o gpurun_out/prof_pairs_r1g python profiles/exp/pairs_time.py > gpurun_out/ncu_pg.log 2>&1
ls -la gpurun_out
