#!/bin/bash
# velocity gauge, every step observed: the observed state is STORED by the slab kernel and reduced by k_observe on the side branch
python -m pytest tests -m gpu -x -q -k "fused_observation or datastores or observ or mesh_api or bench_shapes" 2>&1 | tail -3
python tools/obs_probe.py VEL 512
ION_SLAB_OBS_STORE=0 python tools/obs_probe.py VEL 512
