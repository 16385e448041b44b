#!/bin/bash
# r-segments: LU factors staged by the TMA engine (eight clipped cp.async.bulk per CTA) vs 8 cp.async per thread (exp_nolubseg.so)
N=/root/repo/ionization_b200/_lib/exp_nolubseg.so
python tools/c5_probe.py 4096 40
ION_LIB=$N python tools/c5_probe.py 4096 40
python tools/c5_probe.py 4096 40
ION_LIB=$N python tools/c5_probe.py 4096 40
python -m pytest tests -m gpu -x -q -k "segment or shard or bench_shapes or half_warp" 2>&1 | tail -3
