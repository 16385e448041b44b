#!/bin/bash
python -m pytest tests/test_scan_and_devices.py -m gpu -x -q 2>&1 | grep -v "^  " | tail -40
