cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { python bench.py --config $1 --docs $2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms %.3f frac %.4f launches %d' % (d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']), d['config']['kernel_paths'], d['config']['all_status_ok'], d['config']['replicas_converged'])"; }
run c2 1000; run c3 1000; run c4 20000
