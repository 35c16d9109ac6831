cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider > gpurun_out/r2/pt22.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/r2/pt22.log | tail -30
python scripts/r2/stepbench.py --tag "compacted march walk" --breakdown > gpurun_out/r2/stepbench22.log 2>&1; cat gpurun_out/r2/stepbench22.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2/smoke22.log 2>&1; tail -2 gpurun_out/r2/smoke22.log
timeout 600 python bench.py > gpurun_out/r2/bench22.log 2> gpurun_out/r2/bench22.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2/bench22.log
