cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q -k "not c5" 2>&1 | tail -4
timeout 900 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 full', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['ms_per_step_min'], d['config']['all_status_ok'], d['config']['replicas_converged'], d['gpu_launches'], d['config']['kernel_paths'])"
bash tools/prof2.sh $1 c4 100000 merge_logs_warp_kernel
