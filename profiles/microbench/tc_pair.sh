#!/usr/bin/env bash
# CTA-pair dense kernel vs the single-CTA one, and where the time goes (GCRL_TC_DBG bits: 1 no split, 2 no stores,
# 4 one MMA per k step, 8 cta-scope waits on the cross-CTA barriers, 16 no peer wait (wrong results), 32 no L2 prefetch
# of the next activation tile).  One process: the library reads the switches at every launch.
cd "$(dirname "$0")/../.."
timeout 120 python profiles/microbench/tc_pair.py
