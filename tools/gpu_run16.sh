cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "compact" 2>&1 | tail -4
timeout 900 python bench.py --no-extras --no-cpu-baseline --steps 10 2>gpurun_out/b16.err | tail -1 > gpurun_out/bench_r02_c.json; tail -3 gpurun_out/b16.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_c.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'], d['config']['all_status_ok'])
print('e2e',d.get('e2e'))
"
