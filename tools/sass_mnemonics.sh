#!/bin/bash
# Per-object SASS mnemonic counts of the shipped library (Blackwell-native evidence without rebuilding):
#   bash tools/sass_mnemonics.sh > profiles/r2_sass_mnemonics.txt
cd "$(dirname "$0")/../unopticalflow_b200/lib" || exit 1
M="UTCHMMA UTCBAR LDTM STTM UTMALDG SYNCS.ARRIVE SYNCS.PHASECHK LDGSTS FFMA2 FADD2 FMUL2 REDG BAR.SYNC UCGABAR SHFL MUFU"
echo "# cuobjdump -sass <object> | grep -c <mnemonic>, per object of libuof_b200.so (sm_100a, $(date +%F))"
printf "%-22s" object; for m in $M; do printf "%9s" "${m:0:8}"; done; echo
for o in *.o; do
  s=$(cuobjdump -sass "$o")
  printf "%-22s" "$o"
  for m in $M; do printf "%9d" "$(grep -c -- "$m" <<<"$s")"; done
  echo
done
echo
echo "# tcgen05 / TMEM / TMA instructions of cost_volume_tc.o (opt-in tensor-core backward) and cost_volume_tma.o:"
cuobjdump -sass cost_volume_tc.o | grep -oE "(UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTCATOMSWS|SYNCS)[A-Z0-9_.]*" | sort | uniq -c | sort -rn | head -20
cuobjdump -sass cost_volume_tma.o | grep -oE "(UTMALDG|SYNCS|UCGABAR)[A-Z0-9_.]*" | sort | uniq -c | sort -rn | head -10
