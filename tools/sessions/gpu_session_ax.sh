#!/bin/bash
timeout 1500 python -m pytest tests/test_varlen_matrix_gpu.py -q -m gpu 2>&1 | tail -30 | cut -c1-250
