# Final check of the round on the committed head: the whole GPU suite, smoke(), the default bench command line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_gputest.log
python __graft_entry__.py --smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench rc=$?"
tail -3 gpurun_out/final_gputest.log; tail -2 gpurun_out/final_smoke.log; cut -c1-330 gpurun_out/final_bench_default.json
