#!/bin/bash
# Instruction evidence for the built library: per kernel, how many tcgen05 / TMA / TMEM / bulk-copy / cluster-barrier /
# reduction / elect instructions its SASS holds.    bash scripts/sass_evidence.sh > profiles/r02_sass_evidence.txt
echo "# cuobjdump -sass grafp_b200/lib/libgrafp_b200.so: count, kernel, mnemonic (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store,"
echo "# LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, UCGABAR = cluster barrier, LDGSTS = cp.async, ELECT = elect.sync)"
cuobjdump -sass grafp_b200/lib/libgrafp_b200.so 2>/dev/null | awk '
/Function :/ {fn=$3}
{ for (i = 1; i <= NF; i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|LDTM|UTCBAR|UBLKCP|REDG|RED|ELECT|FMNMX3|UCGABAR_ARV|UCGABAR_WAIT|LDGSTS|ATOMG)/) { m=$i; sub(/;.*/, "", m); c[fn " " m]++ } }
END { for (k in c) print c[k], k }' | sort -k2,2 -k3,3 | grep -E "tc2|bwd_cluster|fwd_pipe|bn_fwd_persistent|bn_bwd_persistent|bn_stats|knn_tc_kernel|ntxent|peak_extract|conv1x1"
