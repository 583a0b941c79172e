#!/usr/bin/env python
"""CUDA-event timing of RunningNormalizer.update / normalize on device-resident float32 batches (bench.py's
normaliser_rooflines alone).  Usage: python profiles/time_normalizer.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
for k, v in bench.normaliser_rooflines(0, flush, torch.cuda.current_stream(dev), float(peaks.get("hbm_gbs", 6650.0))).items():
    print(f"{k:32s} {v['ms_per_launch'] * 1e3:9.2f} us  {v['achieved']:8.1f} GB/s  {100 * v['frac']:5.1f} % of HBM")
