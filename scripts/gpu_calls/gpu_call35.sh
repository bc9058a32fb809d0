#!/bin/bash
timeout 300 python scripts/bench_dgrad_phases.py 2>&1 | tail -10
