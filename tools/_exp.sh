#!/bin/bash
echo "== default routing"; python tools/latency.py 2>&1 | tail -4
echo "== generic (CTA per frame)"; POLAR_B200_FORCE_GENERIC=1 python tools/latency.py 2>&1 | tail -4
echo "== path_warp"; POLAR_B200_FORCE_GENERIC=2 python tools/latency.py 2>&1 | tail -4
