python -m pytest tests/test_ddpg_gpu.py tests/test_td3_gpu.py tests/test_tc_gemm_gpu.py tests/test_dp_gpu.py -m gpu -q --tb=short -k "large_batch or tensor_core or tc_ or odd_shapes or world1 or emulated" > gpurun_out/r2p_pytest.log 2>&1; grep -n "FAILED\|passed\|failed\|Error" gpurun_out/r2p_pytest.log | cut -c1-200 | tail -8
python - <<'PY'
import sys, json, subprocess
PY
python bench.py --steps 50 --warmup 5 --no-cpu --no-big-buffer > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
for k,v in d['rooflines'].items():
    if 'dense' in k or 'gemm' in k: print(k, round(v['ms_per_launch']*1e3,2),'us', round(v['achieved'],1))
print({k:round(v['ms_per_step'],4) for k,v in d['sweep'].items()}); print(d['configs']['ddpg_pickplace_global_B65536']['ms_per_step'])"
