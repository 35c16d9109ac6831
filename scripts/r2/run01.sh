cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt01.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2/pt01.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke01.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2/smoke01.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench01.log 2> gpurun_out/r2/bench01.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2/bench01.err
python scripts/bench_summary.py gpurun_out/r2/bench01.log
