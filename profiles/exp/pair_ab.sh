# A/B of the frame-pair layout of k_rmsd_quad (same box, same call): GROAN_EXP_NO_PAIRS=1 selects one frame per CTA
L=groan_rs_b200
for i in 1 2; do
  timeout 100 python profiles/exp/quad_time.py $L/libquad_head.so 2>&1 | tail -1
  GROAN_EXP_NO_PAIRS=1 timeout 100 python profiles/exp/quad_time.py $L/libgroan_gpu.so 2>&1 | tail -1
  timeout 100 python profiles/exp/quad_time.py $L/libgroan_gpu.so 2>&1 | tail -1
  GROAN_EXP_NO_PAIRS=1 timeout 100 python profiles/exp/quad_time.py $L/libquad_regs.so 2>&1 | tail -1
  timeout 100 python profiles/exp/quad_time.py $L/libquad_regs.so 2>&1 | tail -1
done
