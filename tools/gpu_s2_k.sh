#!/bin/bash
cd "$(dirname "$0")/.."
for d in 64 68; do DISPU_DENSE_DBG=$d timeout 60 python tools/conv1_dbg.py 2>&1 | tail -1; done
