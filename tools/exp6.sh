cd $GRAFT_REPO_ROOT
PT_TMA=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "1 0" "0 0" "1 1" "0 1"; do
  set -- $v
  echo "PT_PREFETCH=$1 PT_TMA=$2"
  PT_PREFETCH=$1 PT_TMA=$2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k "regex:kernel<.int.512>" -s 3 -c 1 --csv python bench.py --config c2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep -E "gpu__time|dram__bytes" | awk -F'","' '{print "   ncu", $(NF-2), $(NF-1), $NF}'
  for C in "c2 1000" "c3 1000" "c4 20000"; do set -- $v $C; PT_PREFETCH=$1 PT_TMA=$2 python bench.py --config $3 --docs $4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   $3 ms %.3f lone %.3f' % (d['ms_per_step'], d['roofline']['launch_ms']), d['config']['all_status_ok'], d['config']['replicas_converged'])"; done
done
