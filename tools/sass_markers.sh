#!/bin/sh
# Counts the SASS mnemonics that prove tcgen05 / TMEM / TMA use, per kernel of the shipped library (B200_PROFILING.md).
# Output: profiles/r02_sass_markers.txt
cd "$(dirname "$0")/.."
LIB=superviseddescent_b200/lib/libsd_b200.so
cuobjdump -sass "$LIB" | awk '
  /Function :/ { name=$3; next }
  /UTCHMMA|UTCQMMA|UTCOMMA/ { mma[name]++ }
  /LDTM|STTM/              { tm[name]++ }
  /UTMALDG/                { ldg[name]++ }
  /UTMASTG|UTMAREDG/       { stg[name]++ }
  /UTCBAR|SYNCS/           { bar[name]++ }
  /\/\*[0-9a-f]+\*\//      { n[name]++ }
  END { printf "%-110s %8s %8s %6s %8s %9s %7s\n", "kernel (mangled)", "instrs", "UTC*MMA", "LDTM", "UTMALDG", "UTMASTG/RED", "SYNCS";
        for (k in n) printf "%-110s %8d %8d %6d %8d %9d %7d\n", k, n[k], mma[k], tm[k], ldg[k], stg[k], bar[k] }' | sort
