for bn in 256 128; do
GCRL_TC_BN=$bn python bench.py --steps 30 --warmup 5 --no-cpu --no-big-buffer > gpurun_out/r2q_bn$bn.json 2> gpurun_out/r2q_bn$bn.err; python -c "
import json
d=json.load(open('gpurun_out/r2q_bn$bn.json'))
for k,v in d['rooflines'].items():
    if 'dense' in k: print('BN=$bn', k, round(v['ms_per_launch']*1e3,2),'us', round(v['achieved'],1))
print({k:round(v['ms_per_step'],4) for k,v in d['sweep'].items()})"
done
