cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt13.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2/pt13.log | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke13.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke13.log
timeout 600 python bench.py > gpurun_out/r2/bench13.log 2> gpurun_out/r2/bench13.err; echo "bench rc=$?"; python scripts/bench_summary.py gpurun_out/r2/bench13.log
# launch list of the same command (cold-cache, serialised per-launch times: shares must agree, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/r2/launches13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2/ncu_launch13.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_ngp_scatter|k_ngp_encode_pair|k_ngp_mlp_bwd|k_ngp_mlp_fwd_ts|k_march_count|k_distill_rays" -s 12 -c 7 -o gpurun_out/r2/prof_top_r13 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-prefetch > gpurun_out/r2/ncu_full13.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2/*.ncu-rep
