cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt31.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt31.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2/smoke31.log 2>&1; tail -2 gpurun_out/r2/smoke31.log
timeout 600 python bench.py > gpurun_out/r2/bench31.log 2> gpurun_out/r2/bench31.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench31.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); print(d['e2e']); print(d['kernel_breakdown_ms_per_step'])
r=d['roofline']; print({k:r[k] for k in r if k not in ('reduction_rate',)}); print(r.get('reduction_rate'))
PY
