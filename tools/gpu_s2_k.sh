#!/bin/bash
cd "$(dirname "$0")/.."
for d in 0 68; do DISPU_DENSE_DBG=$d timeout 60 python tools/conv1_lean_trace.py 2>&1 | tail -18; done
