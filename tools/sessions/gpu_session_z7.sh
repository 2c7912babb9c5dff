#!/bin/bash
export FA_B200_BWD_PAIR=1
timeout 120 python tools/probe_det.py 2>&1 | tail -12
