#!/bin/bash
H=$PWD/flash_attention_b200
export FA_B200_LIB=$H/libfa_b200_ts.so
timeout 120 python tools/probe_bwd_ts.py nc brief 2>&1 | grep -E "once|^it 1[01]" | cut -c1-400
