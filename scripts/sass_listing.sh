#!/bin/bash
# Writes profiles/<round>_sass_counts.txt: per kernel family of the built objects, how many tcgen05 / TMEM / bulk-copy / cluster
# SASS instructions the sm_100a code holds (cuobjdump -sass | grep -c), plus registers / spills from ptxas -v.
# Usage: scripts/sass_listing.sh r02
set -e
R=${1:-r02}
OUT=profiles/${R}_sass_counts.txt
: > $OUT
for o in build/opz_tc2.o build/opz_gemm.o build/opz_bptt.o build/opz_fp32.o; do
  [ -f $o ] || continue
  echo "== $o (cuobjdump -sass, sm_100a)" >> $OUT
  cuobjdump -sass $o > /tmp/sass_$$.txt
  for pat in UTCHMMA UTCQMMA UTCOMMA "UTCHMMA.2CTA" LDTM STTM UTCBAR UBLKCP UTMALDG SYNCS "BAR.SYNC" MUFU "HFMA2" "FFMA" "UCGABAR" ; do
    printf "  %-14s %8d\n" "$pat" "$(grep -c -- "$pat" /tmp/sass_$$.txt || true)" >> $OUT
  done
  echo "  kernels:" >> $OUT
  grep -o "Function : [A-Za-z0-9_]*" /tmp/sass_$$.txt | sed 's/Function : /    /' | c++filt | cut -c1-200 >> $OUT
  rm -f /tmp/sass_$$.txt
done
echo "== ptxas -v (registers, spills)" >> $OUT
for l in build/opz_tc2.ptxas.log build/opz_gemm.ptxas.log build/opz_bptt.ptxas.log build/opz_fp32.ptxas.log; do
  [ -f $l ] || continue
  echo "-- $l" >> $OUT
  grep -A2 "Compiling entry function" $l | grep -v "^--" | sed 's/ptxas info    : //' | paste - - - | c++filt | cut -c1-400 >> $OUT
done
echo wrote $OUT
