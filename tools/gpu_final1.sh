cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== full gpu test suite"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== full-size ncu capture of the warp kernel"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:merge_logs_warp_kernel" -s 3 -c 1 -o gpurun_out/prof_r02_final_c4 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/prof_r02_final_c4.log 2>&1
tail -1 gpurun_out/prof_r02_final_c4.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02_final_c4.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
tail -12 gpurun_out/launches_r02_final_c4.csv | cut -c1-220
