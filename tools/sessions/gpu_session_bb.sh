#!/bin/bash
for t in "" 1 2; do
  echo "== FA_B200_FWD_TILES='$t'"
  FA_B200_FWD_TILES=$t timeout 120 python tools/probe_short.py 2>&1 | tail -8
done
