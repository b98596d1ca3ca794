#!/bin/bash
# profiles/r2_sass_cmac.txt: what the built library holds (cuobjdump; runs on the CPU box)
SO=hisstools_library_b200/lib/libhisstools_b200.so
{
echo "# SASS evidence: cuobjdump -sass of $SO ($(date -u +%F)); built by hisstools_library_b200/build.py for sm_100a only"
echo "# embedded ELF images:"; cuobjdump -lelf $SO
echo; echo "# linked libraries (no cuFFT / cuBLAS / NCCL):"; ldd $SO | awk '{print "#   "$1}'
echo; echo "# mnemonic counts over the whole library (UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 = fma.rn.f32x2, ELECT = elect.sync)"
cuobjdump -sass $SO 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | grep -E "UBLKCP|SYNCS|FFMA2|ELECT|UTMA|HMMA|LDGSTS|UCGABAR|MEMBAR"
echo; echo "# k_cmac_tma<float,1,8> (single-hop tail launch): the bulk copies and mbarrier waits of the 3-stage ring"
cuobjdump -sass -fun '_ZN2hb10k_cmac_tmaIfLi1ELi8EEEvNS_4GeomENS_5RangeEPKNS_5VecOfIT_E4typeES8_PS6_i' $SO 2>/dev/null | grep -E "UBLKCP|SYNCS|BAR.SYNC" | cut -c9-110
echo; echo "# k_cmac_mh2<4,8,32,1> (4 hops per pass of the IR spectra, packed FP32): per work item 128 FFMA2 + 12 LDS.128;"
echo "# operand modifiers do the broadcast (R.F32), the swap (.LO_HI) and the one negated half (-... .NP) of the complex product"
cuobjdump -sass -fun '_ZN2hb10k_cmac_mh2ILi4ELi8ELi32ELi1EEEvNS_4GeomENS_5RangeEPK6float4S5_PS3_im' $SO 2>/dev/null | grep -E "FFMA2|LDS|UBLKCP|SYNCS|ELECT" | cut -c9-120 | head -70
} > profiles/r2_sass_cmac.txt
wc -l profiles/r2_sass_cmac.txt
