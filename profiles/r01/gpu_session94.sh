#!/bin/bash
# where does the per-tile time of the tcgen05 update go?  Timing-only variants of the epilogue's read-modify-write
set -x
mkdir -p gpurun_out
cp draupnir_asr_b200/libdraupnir_b200.so /tmp/keep.so
for v in base NOPREF NOLOAD NOSTORE NORMW base; do
  cp draupnir_asr_b200/ab_$v.so draupnir_asr_b200/libdraupnir_b200.so
  touch draupnir_asr_b200/libdraupnir_b200.so
  echo "== $v"
  timeout 120 python profiles/trace_ozaki.py 256 2>&1 | head -1
  timeout 120 python profiles/trace_ozaki.py 128 2>&1 | head -1
done
cp /tmp/keep.so draupnir_asr_b200/libdraupnir_b200.so
