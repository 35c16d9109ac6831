cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt35.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt35.log | tail -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2/smoke35.log 2>&1; tail -2 gpurun_out/r2/smoke35.log
timeout 600 python bench.py > gpurun_out/r2/bench35.log 2> gpurun_out/r2/bench35.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench35.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); print(d['e2e']); print(d['kernel_breakdown_ms_per_step'])
r=d['roofline']; print({k:r[k] for k in ('kernel','launch_ms','achieved','frac','traffic')}); print({k:v for k,v in r.get('reduction_rate',{}).items() if k!='note'})
PY
timeout 300 python scripts/config23_bench.py > gpurun_out/r2/config23_35.log 2>&1; tail -3 gpurun_out/r2/config23_35.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 220 --csv --log-file gpurun_out/r2/launches35.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2/ncu_launch35.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ngp_scatter|k_ngp_encode_pair|k_ngp_mlp_bwd|k_ngp_mlp_fwd_ts|k_march_walk|k_march_write_samples|k_distill_rays" -s 14 -c 8 -o gpurun_out/r2/prof_top_r35 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-prefetch > gpurun_out/r2/ncu_full35.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2/prof_top_r35.ncu-rep
