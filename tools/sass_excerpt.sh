#!/bin/bash
# SASS evidence for the Blackwell-specific instructions of the kernels (no GPU needed):
#   profiles/r02_sass_k1_k3.txt = opcode histograms + the lines with UBLKCP / SYNCS / LDTM / STTM / UTC*
set -e
OBJ=needle_b200/_obj
OUT=${1:-profiles/r02_sass_k1_k3.txt}
{
echo "# cuobjdump -sass of needle_b200/_obj/*.o (sm_100a), built by needle_b200/build.py"
echo "# PTX -> SASS: cp.async.bulk -> UBLKCP, mbarrier expect_tx/try_wait -> SYNCS, tcgen05.ld/st -> LDTM/STTM, tcgen05.alloc/dealloc -> UTCATOMSWS"
for spec in "fingerprint.o:_ZN5nb20023fp_fft_chroma_tm_kernelILi16ELi4EEEvNS_6K1ArgsE:K1 fp_fft_chroma_tm_kernel<16,4> (default)" \
            "match.o:_ZN5nb20017match_fast_kernelILb1EEEvNS_9MatchArgsE:K3 match_fast_kernel<adaptive> (default)" \
            "match.o:_ZN5nb20017match_fast_kernelILb0EEEvNS_9MatchArgsE:K3 match_fast_kernel<dense>" \
            "match.o:_ZN5nb20012match_kernelENS_9MatchArgsE:K3 match_kernel (general)"; do
  o=${spec%%:*}; rest=${spec#*:}; fn=${rest%%:*}; name=${rest#*:}
  echo; echo "== $name"
  cuobjdump -sass -fun "$fn" $OBJ/$o > /tmp/sass_x.txt
  echo "-- static opcode histogram (top 24)"
  grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_x.txt | awk '{for(i=2;i<=NF;i++){ if ($i !~ /^@/) {print $i; break}}}' | sed 's/;//; s/\..*//' | sort | uniq -c | sort -rn | head -24
  echo "-- TMA / mbarrier / tensor-memory instructions"
  grep -E 'UBLKCP|UTMALDG|SYNCS|LDTM|STTM|UTCATOMSWS|UTCBAR|FENCE\.VIEW' /tmp/sass_x.txt | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' | cut -c1-120
done
} > $OUT
wc -l $OUT
