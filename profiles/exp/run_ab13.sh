mkdir -p gpurun_out
{
for lib in groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_dyn.so groan_rs_b200/libquad_pf4.so groan_rs_b200/libquad_dynpf4.so groan_rs_b200/libquad_pf8.so groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_dyn.so; do
  timeout 150 python profiles/exp/quad_time.py "$lib" 2>&1 | tail -1
done
} > gpurun_out/quad_ab13.txt 2>&1
cat gpurun_out/quad_ab13.txt
