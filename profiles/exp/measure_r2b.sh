# second measurement pass of round 2: the tests added after the first, smoke(), the ncu --set full capture of the FUSED kernel
# (k_rmsd_quad<1, 1>: the fifth launch is inside the timed region of the same command line, run first without ncu)
python -m pytest tests -m gpu -x -q -k "box_spanning or fallback or second_tier or cpp_mirror or fit_short" > gpurun_out/r2_gputest_g.log 2>&1; tail -5 gpurun_out/r2_gputest_g.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_rmsd_quad<1, 1>" -s 4 -c 1 -o gpurun_out/r2_dominant -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
python bench.py --steps 100 --no-cpu --no-e2e > gpurun_out/r2_bench_extras.json 2> gpurun_out/r2_bench_extras.err; echo "bench rc=$?"
