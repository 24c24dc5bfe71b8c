mkdir -p gpurun_out
{
for lib in groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_nb1.so groan_rs_b200/libquad_nb3.so groan_rs_b200/libquad_nb4.so groan_rs_b200/libgroan_gpu.so; do
  timeout 100 python profiles/exp/quad_time.py "$lib" 2>&1 | tail -1
done
FRAME0=37 timeout 100 python profiles/exp/quad_time.py groan_rs_b200/libgroan_gpu.so 2>&1 | tail -1
FRAME0=74 timeout 100 python profiles/exp/quad_time.py groan_rs_b200/libgroan_gpu.so 2>&1 | tail -1
} > gpurun_out/quad_ab26.txt 2>&1
cat gpurun_out/quad_ab26.txt
