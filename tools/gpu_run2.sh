cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== round2 tests"
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -15
for mode in "" "0"; do
echo "== bench c4 30K docs: PT_WARP=$mode"
PT_WARP=$mode timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['launch_ms'], d['config']['kernel_paths'], d['clocks'])"
done
for pf in 1 2 3; do
echo "== bench c4 30K docs: warp, PT_PREFETCH=$pf"
PT_PREFETCH=$pf timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"
done
for w in "2048:4:12:4" "2048:8:12:2" "2048:6:11:3" "2048:4:11:5" "2048:4:14:4"; do
echo "== bench c4 30K docs: PT_WARP=$w"
PT_WARP=$w timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['config']['kernel_paths'])"
done
bash tools/prof2.sh r02_w1_c4 c4 20000 merge_logs_warp_kernel
