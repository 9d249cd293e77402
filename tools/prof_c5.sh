cd $GRAFT_REPO_ROOT
# full ncu capture of the CTA-per-log kernel on a c5 sample (docs = $1), plus the launch list and an un-profiled timing
python bench.py --config c5 --docs $1 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-weak > gpurun_out/c5_bench_$1.json 2> gpurun_out/c5_bench_$1.err
tail -c 1500 gpurun_out/c5_bench_$1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c5.csv python bench.py --config c5 --docs $1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-weak > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:merge_logs_kernel" -s 6 -c 4 -o gpurun_out/prof_c5 python bench.py --config c5 --docs $1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-weak > gpurun_out/prof_c5.log 2>&1
tail -2 gpurun_out/prof_c5.log | cut -c1-200
