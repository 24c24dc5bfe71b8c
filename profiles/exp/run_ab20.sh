mkdir -p gpurun_out
{
for lib in groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_async.so groan_rs_b200/libgroan_gpu.so groan_rs_b200/libquad_async.so; do
  timeout 100 python profiles/exp/quad_time.py "$lib" 2>&1 | tail -1
done
REPS=300 timeout 100 python profiles/exp/quad_time.py groan_rs_b200/libgroan_gpu.so 2>&1 | tail -1
REPS=300 timeout 100 python profiles/exp/quad_time.py groan_rs_b200/libquad_async.so 2>&1 | tail -1
} > gpurun_out/quad_ab20.txt 2>&1
cat gpurun_out/quad_ab20.txt
