cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_tensorf.py tests/test_gpu_parity.py -m gpu -q --timeout 600 --no-header -p no:cacheprovider -k "ffmlp or tensor or linear or colour or distillation_with" > gpurun_out/r2/pt28.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt28.log | tail -30
timeout 300 python scripts/r2/wide_micro.py > gpurun_out/r2/wide_micro28.log 2>&1; cat gpurun_out/r2/wide_micro28.log
timeout 600 python scripts/config5_bench.py > gpurun_out/r2/config5_28.log 2>&1; tail -1 gpurun_out/r2/config5_28.log
