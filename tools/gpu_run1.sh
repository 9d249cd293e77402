cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== parity (both kernel configurations)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adversarial.py -m gpu -x -q 2>&1 | tail -15
echo "== round2"
timeout 1500 python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -15
echo "== rest"
timeout 1500 python -m pytest tests/test_gpu_workloads.py tests/test_gpu_c_abi.py tests/test_facade.py tests/test_multiprocess.py -m gpu -x -q 2>&1 | tail -8
echo "== bench c4 30K docs: warp bin"
timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['config']['kernel_paths'], d['config']['all_status_ok'], d['config']['replicas_converged'])"
echo "== bench c4 30K docs: block only"
PT_WARP=0 timeout 600 python bench.py --config c4 --docs 30000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['config']['kernel_paths'])"
