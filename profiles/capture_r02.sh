#!/usr/bin/env bash
# Round-2 ncu captures (run on a B200 through gpurun; outputs land in gpurun_out/, summaries are copied to profiles/).
#   1. launch list of the update at B = 256 (gpu__time_duration per launch)
#   2. --set full of the row-slab kernels + the weight-gradient and optimiser kernels
#   3. --set full of the sampler at B = 65536 and 1M, and of the normaliser kernels at n = 1M
set -u
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_B256.csv python profiles/prof_driver.py 256 45 > gpurun_out/ncu_r02_l.log 2>&1
$NCU --set full --import-source on -k regex:"fused_critic|fused_actor|wgrad_tile|adam_kernel" -s 32 -c 8 -o gpurun_out/r02_update_B256 -f python profiles/prof_driver.py 256 12 > gpurun_out/ncu_r02_u.log 2>&1
$NCU --set full -k regex:her_sample -s 2 -c 1 -o gpurun_out/r02_sampler_B64k -f python profiles/prof_sampler.py 65536 > gpurun_out/ncu_r02_s1.log 2>&1
$NCU --set full -k regex:her_sample -s 2 -c 1 -o gpurun_out/r02_sampler_B1M -f python profiles/prof_sampler.py 1048576 > gpurun_out/ncu_r02_s2.log 2>&1
$NCU --set full -k regex:"norm_partial|norm_merge|norm_apply" -s 6 -c 3 -o gpurun_out/r02_normalizer_n1M -f python profiles/prof_normalizer.py 1000000 > gpurun_out/ncu_r02_n.log 2>&1
for f in r02_update_B256 r02_sampler_B64k r02_sampler_B1M r02_normalizer_n1M; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r02_* | head -20
