cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt39.log 2>&1
echo "== all gpu tests rc=$?"; tail -25 gpurun_out/pt39.log
