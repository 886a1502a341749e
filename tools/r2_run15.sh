#!/bin/bash
# run 15: warm per-launch-class device times of the batched depth forward
set -x
timeout 300 python tools/depth_spans.py vitb 4 5 2>&1 | tail -14
timeout 300 python tools/depth_spans.py vitl 4 3 2>&1 | tail -14
