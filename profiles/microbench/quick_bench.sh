# usage: quick_bench.sh TAG [extra bench args]  -- one short bench line + the headline numbers
tag=$1; shift
python bench.py --steps 300 --warmup 10 --no-sweep --no-cpu "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err || tail -5 gpurun_out/$tag.err
python -c "
import json
d=json.load(open('gpurun_out/$tag.json')); print('$tag', 'flush', round(d['ms_per_step'],4), 'b2b', round(d['back_to_back']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'critic_k', round(d['roofline']['ms_per_launch'],4), d['clocks']['sm_mhz'])"
