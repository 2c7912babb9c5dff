#!/bin/bash
H=$PWD/flash_attention_b200
export FA_B200_LIB=$H/libfa_b200_ts.so
echo "== non-causal"; timeout 120 python tools/probe_bwd_ts.py nc brief 2>&1 | grep "^it" | head -6 | cut -c1-60
echo "== causal"; timeout 120 python tools/probe_bwd_ts.py causal brief 2>&1 | grep "^it" | head -6 | cut -c1-60
