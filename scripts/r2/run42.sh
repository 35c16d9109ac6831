cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python bench.py > gpurun_out/r2/bench42.log 2> gpurun_out/r2/bench42.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench42.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); print(d['e2e']); print(d['kernel_breakdown_ms_per_step'])
r=d['roofline']; print({k:r[k] for k in ('kernel','launch_ms','achieved','frac','traffic','share_of_step')}); print({k:v for k,v in r.get('reduction_rate',{}).items() if k!='note'})
print(d['step_hbm_roofline']['frac'], d['cpu_baseline']['value'])
PY
python scripts/r2/stepbench.py --tag "final" --breakdown 2>&1 | tail -1 | cut -c1-330
