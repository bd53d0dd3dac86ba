#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q -k "density_region or symmetric_kernel" 2>&1 | tail -6
