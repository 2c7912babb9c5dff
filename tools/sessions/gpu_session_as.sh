#!/bin/bash
timeout 1200 python tools/stress_fwd.py 5 80 2>&1 | grep -E "FAIL|stress done|rror" | cut -c1-330
