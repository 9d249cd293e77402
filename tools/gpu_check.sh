cd $GRAFT_REPO_ROOT
# the whole GPU suite, then a short c5 run (spill path) and the smoke entry
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --config c5 --docs 64 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['all_status_ok'], d['config']['replicas_converged'], d['config']['kernel_paths'])"
python -c "import __graft_entry__ as g; g.smoke()"
