cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { PT_BINS=$1 python bench.py --config $2 --docs $3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '$2', 'ms %.3f frac %.4f' % (d['ms_per_step'], d['roofline']['frac']), d['config']['kernel_paths'], d['config']['all_status_ok'], d['config']['replicas_converged'])"; }
D="1536:128:31:7,4096:256:74:3,6144:512:112:2"
run "$D,0:1024:226:1" c2 1000
run "$D,0:512:112:2" c2 1000
run "$D,0:1024:226:1" c3 1000
run "$D,0:1024:226:1" c4 20000
