mkdir -p gpurun_out
{
for lib in groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_s1.so groan_rs_b200/libquad_s2.so groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_s1.so groan_rs_b200/libquad_s2.so; do
  timeout 100 python profiles/exp/quad_time.py "$lib" 2>&1 | tail -1
done
} > gpurun_out/quad_ab27.txt 2>&1
cat gpurun_out/quad_ab27.txt
