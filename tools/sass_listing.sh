#!/usr/bin/env bash
# Committed SASS evidence (north_star: "plus a committed SASS listing"): the persistent kernels of the fp64 engine,
# encodings stripped.  Look for UBLKCP (cp.async.bulk = TMA), SYNCS.* (mbarrier), LDG.E.128 / STG.E.128 streaming of
# B^-1, DFMA, and — in the sharded kernel — LDG/STG .STRONG.SYS on the peer-mapped mailboxes.
set -euo pipefail
LIB=simplex_method_gpu_b200/libb200lp.so
OUT=profiles
R=${ROUND:-r02}
for F in _ZN6b200lp18simplex_persistentIdLi4ELb0EEEvNS_3DevIT_EE:${R}_sass_simplex_persistent_f64_wc4 \
         _ZN6b200lp18simplex_persistentIdLi8ELb1EEEvNS_3DevIT_EE:${R}_sass_simplex_persistent_steepest_edge_f64_wc8 \
         _ZN6b200lp26simplex_persistent_shardedIdLi8EEEvNS_3DevIT_EE:${R}_sass_simplex_persistent_sharded_f64_wc8 \
         _ZN6b200lp16simplex_residentIdEEvNS_3DevIT_EE:${R}_sass_simplex_resident_f64; do
	SYM=${F%%:*}; NAME=${F##*:}
	cuobjdump -sass -fun "$SYM" "$LIB" 2>/dev/null \
		| grep -E "Function :|^\s+/\*[0-9a-f]{4,}\*/" \
		| sed -E 's@^\s+/\*([0-9a-f]{4,})\*/\s+@\1  @; s@\s*/\* 0x[0-9a-f]+ \*/\s*$@@; s@\s+;$@ ;@' > "$OUT/$NAME.txt"
	echo "$NAME: $(wc -l < "$OUT/$NAME.txt") lines; opcode histogram:" 
	awk 'NR>1 {print $2}' "$OUT/$NAME.txt" | sed 's/\..*//' | sort | uniq -c | sort -rn | head -12 | tr '\n' ';'; echo
	(grep -cE "UBLKCP" "$OUT/$NAME.txt" || true) | sed 's/^/  UBLKCP (TMA bulk copy) instructions: /'
done
