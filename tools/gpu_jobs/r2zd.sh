set -x
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "routing or error_counters or staging or survives" 2>&1 | tail -3
