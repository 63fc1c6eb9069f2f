#!/bin/bash
cd /root/repo
for v in "$@"; do
  echo "=== variant: ${v}"
  if [ "$v" = "default" ]; then timeout 120 python scripts/rowops_ab.py 2>&1 | tail -3
  else DIG_B200_LIB=libdig_b200_${v}.so timeout 120 python scripts/rowops_ab.py 2>&1 | tail -3; fi
done
