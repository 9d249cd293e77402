cd $GRAFT_REPO_ROOT
# usage: prof.sh <tag> <config> <docs> <block>   — full ncu capture of one launch of merge_logs_kernel<block> + launch list
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:kernel<.int.$4>" -s 3 -c 1 -o gpurun_out/prof_$1 python bench.py --config $2 --docs $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_$1.log 2>&1
tail -2 gpurun_out/prof_$1.log | cut -c1-100
ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches_$1.csv python bench.py --config $2 --docs $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
