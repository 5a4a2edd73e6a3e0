timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
bash run_sweep.sh
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep_*.json')):
    try:
        j = json.loads(open(f).read()); print(f.split('sweep_')[1][:-5].ljust(12), round(j['value'],1), round(j['roofline']['frac'],3), {k: round(v,1) for k,v in j['roofline']['stage_ms_rank0'].items()})
    except Exception as e: print(f, 'failed', e)
PY
