cd $GRAFT_REPO_ROOT
for pf in 1 0; do
  echo "PT_PREFETCH=$pf"
  PT_PREFETCH=$pf ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k "regex:kernel<.int.512>" -s 3 -c 2 --csv python bench.py --config c2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep -E "gpu__time|dram__bytes" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
  PT_PREFETCH=$pf python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms %.3f lone %.3f' % (d['ms_per_step'], d['roofline']['launch_ms']))"
done
