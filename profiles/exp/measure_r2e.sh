# Per-warp cp.async ring in the RMSD kernels: the whole GPU suite, same-call A/B against the TMA-ring build, a short bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2e_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_gputest.log
{
for lib in groan_rs_b200/libquad_tma.so groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_tma.so groan_rs_b200/libgroan_gpu.so; do
  timeout 100 python profiles/exp/quad_time.py "$lib" 2>&1 | tail -1
done
} > gpurun_out/quad_ab22.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2e_gputest.log; cat gpurun_out/quad_ab22.txt; cut -c1-900 gpurun_out/r2e_bench.json
