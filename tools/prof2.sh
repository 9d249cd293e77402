cd $GRAFT_REPO_ROOT
# usage: prof2.sh <tag> <config> <docs> <kernel-regex>   — full ncu capture of one launch of the matching kernel + launch list
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$4" -s 3 -c 1 -o gpurun_out/prof_$1 python bench.py --config $2 --docs $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/prof_$1.log 2>&1
tail -2 gpurun_out/prof_$1.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$1.csv python bench.py --config $2 --docs $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
