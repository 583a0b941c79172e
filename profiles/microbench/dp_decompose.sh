# N=2 decomposition of the data-parallel overhead
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 5 --no-sweep --no-cpu $EXTRA 2> gpurun_out/$tag.err | grep '^{' > gpurun_out/$tag.json
python -c "
import json
d=json.load(open('gpurun_out/$tag.json')); print('$tag', 'flush', round(d['ms_per_step'],4), 'b2b', round(d['back_to_back']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['per_step_ms']['slowest'], d['per_step_ms']['median_ms'], d['dp_parity'])"; }
run dpd2_p2p A=1
run dpd2_nowait GCRL_P2P_DEBUG=1
run dpd2_nopdl GCRL_NO_PDL=1
