# A/B timing of library variants through bench.py (tuning experiments): bash profiles/exp/ab.sh libgroan_gpu [libexp1 ...]
for v in "${@:-libgroan_gpu}"; do
GROAN_GPU_LIB=$PWD/groan_rs_b200/$v.so timeout 120 python bench.py --steps 100 --warmup 3 --no-extras --no-cpu --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
python - <<P
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["ms_per_step"],4), [round(v["ms"],4) for v in d["roofline"]["ops"].values()], d["clocks"]["sm_mhz"], d["roofline"]["fallback_frames"])
P
done
