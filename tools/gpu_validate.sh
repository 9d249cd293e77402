cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/sanitize2.sh > gpurun_out/sanitize_final.log 2>&1; grep -c "Error\|hazard" gpurun_out/sanitize_final.log; tail -4 gpurun_out/sanitize_final.log
timeout 1500 python bench.py > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
tail -c 300 gpurun_out/bench_$1.json
