cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_workloads.py tests/test_gpu_round2.py tests/test_gpu_adversarial.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_workloads.py -m gpu -x -q -k "generated_workloads_match_oracle and c2-24-2500" 2>&1 | grep -E "SUMMARY|passed|failed|Error" | tail -4
timeout 300 python bench.py --config c2 --no-extras --no-cpu-baseline --no-e2e --steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['config']['all_status_ok'], d['config']['replicas_converged'], d['config']['kernel_paths'])"
