#!/bin/bash
# SASS evidence for the TMA / mbarrier / dp4a claims of DESIGN.md: per kernel of libmcaller_b200.so, how many bulk-copy
# (UBLKCP), mbarrier (SYNCS.*) and dot-product (IDP.4A) instructions it holds, plus the first occurrences with their
# addresses.  usage: tools/sass_excerpt.sh > profiles/r2_k_scan_sass_excerpt.txt
LIB=${1:-mcaller_b200/libmcaller_b200.so}
echo "# cuobjdump -sass $LIB  (sm_100a cubin; $(date -u +%Y-%m-%d))"
cuobjdump -sass "$LIB" | awk '
/Function : / { fn=$3; next }
/UBLKCP|SYNCS|IDP\.4A|DFMA|MUFU\.RCP64H|LDGSTS|ATOMS|RED\./ {
    op=$0; sub(/^[ \t]*\/\*[0-9a-f]+\*\/[ \t]*/, "", op); split(op, t, " "); m=t[1]; if (m ~ /^@/) m=t[2];
    key=fn "\t" m; cnt[key]++; if (!(key in first)) first[key]=$0
}
END {
    for (k in cnt) print cnt[k] "\t" k
}' | sort -k2,2 -k1,1nr | awk -F'\t' '{ printf "%-8s %-34s %s\n", $1, $3, $2 }' | c++filt 2>/dev/null | sed 's/(anonymous namespace):://' | cut -c1-200
echo
echo "# first occurrences in k_scan (three of each kind)"
cuobjdump -sass "$LIB" | awk '/Function : /{f=($0 ~ /6k_scan/); if (f) print} f && /UBLKCP|SYNCS\.ARRIVE|SYNCS\.PHASECHK|IDP\.4A/ {
    k = ($0 ~ /UBLKCP/) ? "a" : ($0 ~ /ARRIVE/) ? "b" : ($0 ~ /PHASECHK/) ? "c" : "d"; if (++n[k] <= 3) print }' | cut -c1-160 | head -40
