cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python bench.py 2>gpurun_out/final_n1.err | tail -1 > gpurun_out/bench_r02_final.json; tail -2 gpurun_out/final_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_final.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'traffic',d['roofline']['traffic'], d['config']['all_status_ok'], d['config']['replicas_converged'])
print('e2e',{k:v for k,v in d.get('e2e',{}).items() if k!='api'})
print('extras',{k:(v['ms_per_step'],v['roofline_frac']) for k,v in d.get('extra_configs',{}).items()})
print('ingest',d.get('ingest',{}).get('mb_per_s'), 'cpu',d.get('cpu_baseline',{}).get('value'))
"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
