cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "level_chunks or fused_backward or training_engines" > gpurun_out/pt32.log 2>&1
echo "rc=$?"; tail -12 gpurun_out/pt32.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench32.log 2>&1; python scripts/bench_summary.py gpurun_out/bench32.log
