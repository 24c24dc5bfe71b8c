# Pair ring (two frames per CTA against one copy of the reference): quick GPU checks and same-call A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "rmsd or fused or quad or center or bench or odd or kabsch or golden" > gpurun_out/r2f_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_gputest.log
{
for lib in groan_rs_b200/libquad_tma.so groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_tma.so groan_rs_b200/libgroan_gpu.so; do
  timeout 100 python profiles/exp/quad_time.py "$lib" 2>&1 | tail -1
done
} > gpurun_out/quad_ab23.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu --no-extras --no-e2e > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2f_gputest.log; cat gpurun_out/quad_ab23.txt; cut -c1-400 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
