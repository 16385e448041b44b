#!/bin/bash
python -m pytest tests -m gpu -x -q -k "shard or peer or cuts or config5 or sharded or devices" 2>&1 | tail -8
