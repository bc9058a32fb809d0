#!/bin/bash
timeout 600 python scripts/debug_graph_streams.py 2>&1 | grep graphs=
CH=32 timeout 600 python scripts/debug_graph_streams.py 2>&1 | grep graphs=
