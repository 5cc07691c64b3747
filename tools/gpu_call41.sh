#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
timeout 120 python -m pytest tests/test_reference_galt.py -x -q -m gpu -s 2>&1 | tail -12
