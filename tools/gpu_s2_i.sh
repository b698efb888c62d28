#!/bin/bash
cd "$(dirname "$0")/.."
timeout 100 python tools/edge_trace.py 2>&1 | tail -4
timeout 150 python -m pytest tests -m gpu -q -x -k "generator or feature or edge or dense_conv" 2>&1 | tail -2
