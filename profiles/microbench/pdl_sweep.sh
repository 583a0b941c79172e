python -m pytest tests/test_dp_gpu.py tests/test_ddpg_gpu.py tests/test_td3_gpu.py -m gpu -x -q 2>&1 | tail -4
for m in 0 15 1 2 4 6 3 5; do
  GCRL_PDL_MASK=$m python bench.py --steps 300 --warmup 10 --no-sweep --no-cpu > gpurun_out/r2b_m$m.json 2> gpurun_out/r2b_m$m.err
  python -c "
import json
d=json.load(open('gpurun_out/r2b_m$m.json')); print('mask $m', round(d['ms_per_step'],4), round(d['back_to_back']['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
