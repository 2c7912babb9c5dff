#!/bin/bash
timeout 1500 python -m pytest tests/test_reference_matrix_gpu.py tests/test_fwd_stress_gpu.py tests/test_bwd_stress_gpu.py -q -m gpu 2>&1 | tail -15 | cut -c1-250
