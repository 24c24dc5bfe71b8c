# A/B timing of library builds (tuning experiments): bash profiles/exp/ab.sh groan_rs_b200/libgroan_gpu.so [groan_rs_b200/libquad_x.so ...]
# Builds of variants: profiles/exp/quad_variants.sh name:"-DKNOB=1"; the timing itself is profiles/exp/quad_time.py (bursts of 50
# launches of the three ring-fed ops on the bench's workload; REPS=300 for the sustained regime, FRAME0 to pick the frames).
for lib in "${@:-groan_rs_b200/libgroan_gpu.so}"; do
  timeout 200 python profiles/exp/quad_time.py "$lib" 2>&1 | tail -1
done
