cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== round2 + parity tests"
timeout 1800 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_adversarial.py -m gpu -x -q 2>&1 | tail -15
run() { timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['kernel_paths'])"; }
echo "== default"; run
for pf in 1 2 3; do echo "== PT_PREFETCH=$pf"; PT_PREFETCH=$pf run; done
for w in "2048:8:6144:4" "2048:16:6656:2" "2048:4:6656:8" "2048:8:7168:4" "2048:6:12288:3" "2048:8:8192:3"; do
echo "== PT_WARP=$w"; PT_WARP=$w run
done
bash tools/prof2.sh r02_w2_c4 c4 20000 merge_logs_warp_kernel
