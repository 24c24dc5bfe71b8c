set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; echo rc=$?
