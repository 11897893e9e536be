#!/bin/bash
for m in dummy nodiag step; do
  echo "=== MODE=$m"
  MODE=$m python tools/e2e_probe.py 2>&1 | grep -E "wall|step[123] |down[012] |up[234] "
done
