cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== parity + workloads"
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adversarial.py tests/test_gpu_workloads.py tests/test_gpu_round2.py -m gpu -x -q -k "not c5" 2>&1 | tail -5
run() { timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"; }
for f in 4 260 516 1028 2052 1540 3844; do echo "== PT_WARP_FLAGS=$f"; PT_WARP_FLAGS=$f run; done
for c in c2 c3; do echo "== $c"; timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['all_status_ok'])"; done
