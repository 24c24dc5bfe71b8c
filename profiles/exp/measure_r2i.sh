mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "tric or center or centre" > gpurun_out/r2i_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_gputest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r2i_gputest.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
for k,v in d.get('extras',{}).items():
    if 'tric' in k or 'center' in k: print(k, v)
print(d['ms_per_step'], d['roofline']['ops']['group_get_center'])
PY
