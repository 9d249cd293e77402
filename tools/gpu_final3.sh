cd $GRAFT_REPO_ROOT
bash tools/prof2.sh $1 c4 100000 merge_logs_warp_kernel
timeout 1500 python bench.py > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
tail -c 600 gpurun_out/bench_$1.json
