#!/bin/bash
# SASS instruction census of every object of libpatchaug_b200.so (proves which kernels use tcgen05 / TMEM / TMA): profiles/r02_sass_census.md
cd "$(dirname "$0")/../patchaugnet_b200/build" || exit 1
out=../../profiles/r02_sass_census.md
{
echo "# SASS census of patchaugnet_b200/build/*.o (cuobjdump -sass | grep -c), sm_100a"
echo
echo "UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / .st (tensor memory), UTMALDG = TMA tensor load,"
echo "UBLKPF = bulk L2 prefetch, SYNCS = mbarrier ops, REDUX = redux.sync, F*2 = packed fp32 pair instructions, MUFU.EX2 = exp2."
echo
echo "| object | UTCHMMA | UTCBAR | LDTM | STTM | UTMALDG | UBLKPF | SYNCS | REDUX | FFMA2/FADD2/FMUL2 | MUFU.EX2 |"
echo "|---|---|---|---|---|---|---|---|---|---|---|"
for f in *.o; do
  cuobjdump -sass "$f" > /tmp/sass_census.txt 2>/dev/null
  c() { grep -c "$1" /tmp/sass_census.txt; }
  echo "| $f | $(c UTCHMMA) | $(c UTCBAR) | $(c LDTM) | $(c STTM) | $(c UTMALDG) | $(c 'UBLKPF\|UTMAPF') | $(c SYNCS) | $(c REDUX) | $(c 'FFMA2\|FADD2\|FMUL2') | $(c MUFU.EX2) |"
done
} > "$out"
cat "$out"
