#!/bin/bash
timeout 1500 python -m pytest tests/test_kvcache_matrix_gpu.py -q -m gpu -x 2>&1 | tail -30 | cut -c1-250
