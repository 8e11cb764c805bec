#!/bin/bash
( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -5
timeout 300 python scripts/gpu_c4_check.py time trace trace_bwd 2>&1 | tail -40
