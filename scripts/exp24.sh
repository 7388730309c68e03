#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for s in c2 2236; do python scripts/sweep2.py $s 12:16; done
