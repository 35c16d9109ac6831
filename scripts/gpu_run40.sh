cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout 300 --no-header -p no:cacheprovider -k "backward or chunks or training_engines or fused" > gpurun_out/pt40.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/pt40.log
timeout 300 python scripts/kbench.py --set 12:1,0,1,0 2>&1 | grep variant | tee gpurun_out/kbench40.log
