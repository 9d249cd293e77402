cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== full gpu test suite"
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== full bench default"
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_r02_b.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_b.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
print('e2e',d.get('e2e'))
print('extras',{k:(v['ms_per_step'],v['roofline_frac']) for k,v in d.get('extra_configs',{}).items()})
print('ingest',d.get('ingest'))
print('cpu',d.get('cpu_baseline'))
"
