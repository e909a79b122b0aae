#!/bin/bash
for dev in 0 1 2 3; do
  echo "== FOLD=1 DEV=$dev"
  CCEDIT_ATTN_FOLD=1 CCEDIT_ATTN_DEV=$dev timeout 300 python tools/dev_attn.py 2>&1 | grep -E "BAD|attn F=34 L=6144|scale=4"
done
echo "== FOLD=0"; CCEDIT_ATTN_FOLD=0 timeout 300 python tools/dev_attn.py 2>&1 | grep -E "BAD|attn F=34 L=6144"
