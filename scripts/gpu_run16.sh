cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/trace_bwd.py > gpurun_out/trace_bwd2.log 2>&1; echo "trace rc=$?"
tail -24 gpurun_out/trace_bwd2.log
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python scripts/kbench.py --rays 262144 2>&1 | grep "^{"
