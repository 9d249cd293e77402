cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== round2 + parity tests"
timeout 1800 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_adversarial.py -m gpu -x -q -k "not c5 and not 32000" 2>&1 | tail -5
run() { timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['kernel_paths'])"; }
echo "== phased 8x4 (default)"; run
echo "== phased 16x2"; PT_WARP=2048:16:6656:2 run
