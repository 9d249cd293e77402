cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== round2 + parity tests"
timeout 1800 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_adversarial.py -m gpu -x -q 2>&1 | tail -8
echo "== parity tests free mode"
PT_WARP_FLAGS=0 timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adversarial.py tests/test_gpu_round2.py -m gpu -x -q -k "not c5 and not 32000" 2>&1 | tail -4
run() { timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['kernel_paths'])"; }
echo "== phased 8x4 (default)"; run
echo "== phased 16x2"; PT_WARP=2048:16:6656:2 run
echo "== phased 12x2 8.7KB"; PT_WARP=2048:12:8960:2 run
echo "== phased 8x4 + prefetch"; PT_WARP_FLAGS=7 run
echo "== free"; PT_WARP_FLAGS=0 run
bash tools/prof2.sh r02_w4_c4 c4 20000 merge_logs_warp_kernel
echo "== full bench default"
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r02_a.json; cat gpurun_out/bench_r02_a.json | cut -c1-3000
