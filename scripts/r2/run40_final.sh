cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt40.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt40.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2/smoke40.log 2>&1; tail -2 gpurun_out/r2/smoke40.log
timeout 600 python bench.py > gpurun_out/r2/bench40.log 2> gpurun_out/r2/bench40.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench40.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); print(d['e2e'])
r=d['roofline']; print({k:r[k] for k in ('kernel','launch_ms','achieved','frac','traffic')}); print({k:v for k,v in r.get('reduction_rate',{}).items() if k!='note'})
PY
for b in 128 512 1024; do S3D_SCATTER_BLOCK=$b python scripts/r2/stepbench.py --tag "scatter block $b" --breakdown 2>&1 | tail -1 | cut -c1-260; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 400
