#!/bin/bash
timeout 300 python scripts/bench_sustained.py 2>&1 | tail -14
