#!/bin/bash
for cfg in "1 2048" "1 2048" "1 1536"; do
  set -- $cfg
  echo "== NT=$1 PIECE_KB=$2"
  MVS_COPY_NT=$1 MVS_COPY_PIECE_KB=$2 python scripts/probe_hostcopy.py | grep staged | tr -d '\n'; echo
done
