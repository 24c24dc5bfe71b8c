# Round-2 closing pass with the per-warp cp.async ring (one gpurun call): the driver's command line for both arms, the ncu launch
# list and a --set full capture of the fused kernel (each ncu run after the same command exited 0 without it).
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2g_bench_reference_arm.json 2> gpurun_out/r2g_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2>> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2g_bench_short.json 2>> gpurun_out/r2g_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2g_ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_rmsd_quad<.bool.1, .int.1>" -s 4 -c 1 -o gpurun_out/r2g_dominant -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2g_ncu_full.log 2>&1
cut -c1-400 gpurun_out/r2g_bench.json; cut -c1-300 gpurun_out/r2g_bench_reference_arm.json; tail -2 gpurun_out/r2g_ncu_full.log
