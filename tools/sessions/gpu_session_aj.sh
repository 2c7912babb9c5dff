#!/bin/bash
H=$PWD/flash_attention_b200
FA_B200_LIB=$H/libfa_b200_fts.so timeout 120 python tools/probe_fwd_ts.py 2>&1 | tail -9 | cut -c1-420
timeout 200 python tools/probe_bwd.py time 2>&1 | grep "time\]"
