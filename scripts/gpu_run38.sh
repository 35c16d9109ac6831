cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 --no-header -p no:cacheprovider -k "grid or golden or reference_outputs or training_engines or field_forward" > gpurun_out/pt38.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/pt38.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench38.log 2>&1; python scripts/bench_summary.py gpurun_out/bench38.log
