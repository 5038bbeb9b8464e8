#!/bin/bash
# Debug build of the C-ABI library WITH the tcgen05 update kernel's phase-accounting / event-trace hooks
# (drp_ozaki_set_prof / drp_ozaki_set_trace).  They are process-global state, so the product library
# (draupnir_asr_b200/libdraupnir_b200.so) does not contain them; profiles/prof_ozaki.py and profiles/trace_ozaki.py load
# this one instead (DRP_LIB=profiles/libdraupnir_b200_debug.so).
set -e
cd "$(dirname "$0")/../draupnir_asr_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DDRP_DEBUG_HOOKS -Xcompiler -fPIC,-fvisibility=hidden \
     -shared -o ../../profiles/libdraupnir_b200_debug.so api.cu factor.cu ozaki.cu ou.cu patristic.cu catlik.cu gru.cu adam.cu sites.cu -lcudart
