#!/bin/bash
# usage: scripts/sass_summary.sh > profiles/r2_sass_summary.txt   -- mnemonic counts of the shipped cubin
cd "$(dirname "$0")/.."
LIB=ncrystal_b200/lib/libncrystal_b200.so
SASS=$(mktemp -p gpurun_out sass.XXXXXX)
cuobjdump -sass $LIB > $SASS
echo "# SASS summary of $LIB (cuobjdump -sass), commit $(git log -1 --format=%h)"
echo "arch: $(cuobjdump -lelf $LIB | tr '\n' ' ')"
echo "kernels (Function :): $(grep -c 'Function :' $SASS)"
for m in UBLKCP SYNCS DFMA DMUL DADD DSETP MUFU.RCP64H HMMA UTCHMMA TCGEN LDG STG ATOMG ATOMS RED LDS STS BAR.SYNC SHFL VOTE MATCH REDUX FFMA; do
  echo "$m: $(grep -c "[ .]$m[ .]" $SASS)"
done
echo
echo "# per kernel: registers / stack / shared (cuobjdump -res-usage)"
cuobjdump -res-usage $LIB 2>/dev/null | awk '/Function/ {f=$2} /REG:/ {print f" "$0}' | c++filt | sed 's/(.*)://; s/  */ /g' | sort -u
rm -f $SASS
