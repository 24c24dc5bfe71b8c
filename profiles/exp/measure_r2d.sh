# Round-2 closing pass on the committed head (one gpurun call): the whole GPU suite, the driver's command line for both arms,
# the ncu launch list and a --set full capture of the final fused kernel (each ncu run after the same command exited 0 without it).
python -m pytest tests -m gpu -q > gpurun_out/r2d_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_gputest.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2d_bench_reference_arm.json 2> gpurun_out/r2d_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2>> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2d_bench_short.json 2>> gpurun_out/r2d_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2d_ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_rmsd_quad<.bool.1, .int.1>" -s 4 -c 1 -o gpurun_out/r2d_dominant -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2d_ncu_full.log 2>&1
tail -3 gpurun_out/r2d_gputest.log; cat gpurun_out/r2d_bench.json | cut -c1-600
