cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/n2.err | tail -1 > gpurun_out/bench_r02_n2.json
tail -5 gpurun_out/n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_n2.json'))
print('value',d['value'],'ms',d['ms_per_step'],'n',d['n_gpus'],d['scaling'],'frac',d['roofline']['frac'])
print('cfg', {k:d['config'][k] for k in ('docs_per_gpu','all_status_ok','replicas_converged','exchange','rank0_cpu_binding','ms_per_step_min','ms_per_step_max')})
print('e2e',d.get('e2e'))
print('weak',d.get('weak'))
"
