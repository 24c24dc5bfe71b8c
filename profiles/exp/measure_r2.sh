# Round-2 measurement pass (one gpurun call): GPU tests, bench line, reference arm, ncu launch list and full capture of the
# dominant kernel of the SAME command line (each ncu run after that command has exited 0 without ncu).
python -m pytest tests -m gpu -q > gpurun_out/r2_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputest.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2_bench_short.json 2>> gpurun_out/r2_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_rmsd_quad<.bool.1, .int.1>" -s 4 -c 1 -o gpurun_out/r2_dominant -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
tail -4 gpurun_out/r2_gputest.log
