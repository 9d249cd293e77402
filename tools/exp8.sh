cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for G in 1 0; do echo PT_GRAPH=$G; for C in "c2 1000" "c3 1000" "c4 20000"; do set -- $C; PT_GRAPH=$G python bench.py --config $1 --docs $2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   $1 ms %.3f lone %.3f frac %.4f launches %d' % (d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['gpu_launches']), d['config']['all_status_ok'], d['config']['replicas_converged'])"; done; done
