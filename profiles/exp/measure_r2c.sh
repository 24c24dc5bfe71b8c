# ncu --set full capture of the FUSED kernel (k_rmsd_quad<1, 1>): the fifth launch is inside the timed region of the same command
# line, which is run first without ncu; then the cell-grid search timings (store path) and its parity tests
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_rmsd_quad<.bool.1, .int.1>" -s 4 -c 1 -o gpurun_out/r2_dominant -f \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
python -m pytest tests -m gpu -x -q -k "pairs_within or cpp_mirror or full_size" 2>&1 | tail -3
python profiles/exp/cells_time.py 2>&1 | tail -6
