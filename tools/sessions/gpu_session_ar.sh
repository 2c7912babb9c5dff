#!/bin/bash
timeout 900 python tools/stress_bwd.py 3 120 2>&1 | grep -E "FAIL|stress done" | cut -c1-220
timeout 900 python tools/stress_bwd.py 4 120 2>&1 | grep -E "FAIL|stress done" | cut -c1-220
