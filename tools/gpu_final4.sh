cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for f in 0 1796; do
PT_WARP_FLAGS=$f timeout 900 python bench.py --no-extras --no-cpu-baseline --no-e2e --steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flags $f: c4 full', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['ms_per_step_min'], d['config']['all_status_ok'])"
done
bash tools/sanitize2.sh > gpurun_out/sanitize_final.log 2>&1; tail -5 gpurun_out/sanitize_final.log
