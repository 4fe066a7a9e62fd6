#!/bin/bash
# A/B of the fused table publish of the sharded loop on one rank's share of ML-20M / 8 (one GPU)
python -m pytest tests/test_em_gpu.py -x -q -k "sharded_engine" 2>&1 | tail -2
MMSBM_SHARD_FUSE=1 python profiles/scripts/rank_compute.py ml20m 8 0 > gpurun_out/fuse1.json
MMSBM_SHARD_FUSE=0 python profiles/scripts/rank_compute.py ml20m 8 0 > gpurun_out/fuse0.json
python profiles/scripts/rank_compute.py netflix 8 3 > gpurun_out/fuse1_netflix.json
tail -n 1 gpurun_out/fuse1.json gpurun_out/fuse0.json gpurun_out/fuse1_netflix.json
