#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
