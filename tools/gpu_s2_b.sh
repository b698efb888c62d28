#!/bin/bash
cd "$(dirname "$0")/.."
python tools/dense_clock_probe.py 32768x2048x256 32768x256x256 2>&1 | tail -8
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
