#!/bin/bash
timeout 600 python -m pytest tests/test_tools.py -x -q 2>&1 | tail -15
