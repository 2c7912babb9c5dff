#!/bin/bash
timeout 900 python tools/stress_bwd.py 1 80 2>&1 | tail -25 | cut -c1-200
FA_B200_BWD_PAIR=0 timeout 600 python tools/stress_bwd.py 2 30 2>&1 | tail -6 | cut -c1-200
