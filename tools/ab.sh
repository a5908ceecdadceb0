#!/bin/bash
# Development aid: time several build variants (tools/build_variant.py) in one gpurun call.
# usage: tools/ab.sh "<counts>" variant...   ("base" = the in-tree library; NOTE: build_variant.py
# rebuilds the in-tree library from the CURRENT sources, so "base" is never an older kernel: build
# the reference point as a variant of its own, e.g. from a git worktree of the old commit)
counts="$1"; shift
for v in "$@"; do
  if [ "$v" = base ]; then lib=cpu-gpu-tfhe_b200/libtfhe_b200.so; else lib=build/variants/$v/libtfhe_b200.so; fi
  echo "== $v"
  TFHE_B200_LIB=$lib timeout 150 python tools/quick_bench.py $counts 2>&1 | grep -v load_keys | tail -6
  rc=${PIPESTATUS[0]}; [ "$rc" != 0 ] && echo "   rc=$rc (124 = hung)"
done
