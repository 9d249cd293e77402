cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== round2 + parity tests (free mode)"
timeout 1800 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_adversarial.py -m gpu -x -q 2>&1 | tail -8
echo "== round2 + parity tests (phased mode)"
PT_PREFETCH=4 timeout 1800 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_adversarial.py -m gpu -x -q -k "not c5 and not 32000" 2>&1 | tail -8
run() { timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['kernel_paths'])"; }
echo "== free default 8x4"; run
echo "== phased 8x4"; PT_PREFETCH=4 run
echo "== phased 16x2"; PT_PREFETCH=4 PT_WARP=2048:16:6656:2 run
echo "== phased 4x8"; PT_PREFETCH=4 PT_WARP=2048:4:6656:8 run
echo "== phased+prefetch marks 16x2"; PT_PREFETCH=5 PT_WARP=2048:16:6656:2 run
echo "== free 16x2"; PT_WARP=2048:16:6656:2 run
PT_PREFETCH=4 PT_WARP=2048:16:6656:2 bash tools/prof2.sh r02_w3_c4 c4 20000 merge_logs_warp_kernel
