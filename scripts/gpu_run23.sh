cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider -k "one_step_ahead" > gpurun_out/pt23.log 2>&1
echo "rc=$?"; tail -30 gpurun_out/pt23.log
timeout 600 python bench.py --no-cpu-baseline --no-roofline --no-prefetch > gpurun_out/bench23_inline.log 2>&1; tail -1 gpurun_out/bench23_inline.log > /dev/null; python scripts/bench_summary.py gpurun_out/bench23_inline.log
timeout 600 python bench.py --no-cpu-baseline --no-roofline > gpurun_out/bench23_ahead.log 2>&1; tail -1 gpurun_out/bench23_ahead.log > /dev/null; python scripts/bench_summary.py gpurun_out/bench23_ahead.log
