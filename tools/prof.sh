cd $GRAFT_REPO_ROOT
# usage: prof.sh <tag> <config> <docs> <kernel-regex>
ncu --set full --clock-control none --import-source on -k "regex:$4" -s 3 -c 1 -o gpurun_out/prof_$1 python bench.py --config $2 --docs $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_$1.log 2>&1
tail -1 gpurun_out/prof_$1.log | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_$1.csv python bench.py --config $2 --docs $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
