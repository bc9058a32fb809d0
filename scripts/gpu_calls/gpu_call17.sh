#!/bin/bash
export ONLY=1
for e in "X=1" "IDEAS_B200_UMMA=0" "IDEAS_AB_NO_PREMOD=1" "IDEAS_OPTS=pmh=0" "IDEAS_AB_NO_CACHE=1" "IDEAS_OPTS=pmh=0,halo=0"; do
  echo "== $e"; env $e timeout 300 python scripts/debug_graph_streams.py 2>&1 | grep graphs=
done
