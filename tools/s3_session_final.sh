#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full gpu tests"; (time timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.log 2>&1); tail -3 gpurun_out/r02_gputests_final.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== reference arm"; python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 300 gpurun_out/r02_bench_reference_arm.json; echo
echo "== full bench"; (time python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err); python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_final.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['levels_0_2'], d['clocks'])
for k in ['train_step','dae_decode','ddec_forward']: print(k, d[k]['value'], d[k].get('roofline',{}).get('frac'))
print('gpu_eager', d['gpu_eager']['sampler_step'], d['gpu_eager'].get('train_step'))
print('cpu', d['cpu_baseline']['value'])
print('optim', d['optim_step']['ms'], d['optim_step']['roofline']['frac'], d['optim_step']['train_step_with_optimizer']['value'])
s=d['secondary']; print('secondary', s.get('value'), json.dumps(s)[:300])
"
