#!/bin/bash
H=$PWD/flash_attention_b200
export FA_B200_LIB=$H/libfa_b200_ts.so
timeout 120 python tools/probe_bwd_spans.py 2>&1 | tail -6
timeout 120 python tools/probe_bwd_spans.py causal 2>&1 | tail -6
