cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pt26.log 2>&1
echo "== all gpu tests rc=$?" >> gpurun_out/summary.txt; tail -6 gpurun_out/pt26.log >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke26.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke26.log >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_r26.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_r26.log >> gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r26.log 2>&1; echo "bench ref rc=$?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_ref_r26.log >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 160 --csv --log-file gpurun_out/launches_r26.csv python bench.py --rays 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch26.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ngp_scatter|k_ngp_encode_pair|k_ngp_mlp_bwd" -s 9 -c 3 -o gpurun_out/prof_top3_r26 python bench.py --rays 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-prefetch > gpurun_out/ncu_full26.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
