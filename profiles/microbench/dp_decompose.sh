# N-GPU comparison of the peer-memory averaging variants (usage: dp_decompose.sh N)
N=${1:-2}
run() { tag=$1; shift; env GCRL_P2P_TIMEOUT_MS=3000 "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 --no-sweep --no-cpu $EXTRA 2> gpurun_out/$tag.err | grep '^{' > gpurun_out/$tag.json
python -c "
import json
d=json.load(open('gpurun_out/$tag.json')); print('$tag', 'flush', round(d['ms_per_step'],4), 'b2b', round(d['back_to_back']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['per_step_ms']['slowest'], d['per_step_ms']['median_ms'], d['dp_parity'])" || tail -5 gpurun_out/$tag.err; }
run dpd3_tile_n$N A=1
run dpd3_launch_n$N GCRL_P2P_TILE_FUSED=0
