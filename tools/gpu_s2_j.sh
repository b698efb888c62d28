#!/bin/bash
cd "$(dirname "$0")/.."
timeout 200 python -m pytest tests -m gpu -q -x -k "lean_and_general or dynamic_range" 2>&1 | tail -15
