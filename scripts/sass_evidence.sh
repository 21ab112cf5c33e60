#!/bin/bash
# Per-kernel SASS evidence of the sm_100a mechanisms the design claims:
#   UBLKCP = TMA bulk copies (cp.async.bulk), SYNCS = mbarrier try_wait/arrive, REDUX = warp reductions (redux.sync),
#   RED.E.ADD.F64 / REDG = fp64 reductions of the scatter formulation, MATCH / VOTE = warp match / ballot (sort, routing).
# usage: scripts/sass_evidence.sh > profiles/r2_sass_counts.txt
set -e
LIB=${1:-easylp_b200/libeasylp_b200.so}
echo "# $(date -u +%FT%TZ)  $(nvcc --version | tail -1)  $LIB"
echo "# kernel | instructions | UBLKCP | SYNCS | REDUX | RED/REDG.ADD.F64 | MATCH | VOTE | BAR.SYNC | MEMBAR"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { if (name != "") print name, n, ub, sy, rx, rd, ma, vo, ba, mb; name=$3; n=ub=sy=rx=rd=ma=vo=ba=mb=0; next }
  /^ *\/\*[0-9a-f]+\*\// { n++;
     if ($0 ~ /UBLKCP/) ub++; if ($0 ~ /SYNCS/) sy++; if ($0 ~ /REDUX/) rx++;
     if ($0 ~ /RED(G)?\.E\.ADD\.F64/) rd++; if ($0 ~ /MATCH/) ma++; if ($0 ~ /VOTE/) vo++; if ($0 ~ /BAR\.SYNC/) ba++; if ($0 ~ /MEMBAR/) mb++ }
  END { if (name != "") print name, n, ub, sy, rx, rd, ma, vo, ba, mb }' | while read name rest; do
    echo "$(echo $name | c++filt | sed 's/(.*//' | cut -c1-110) | $(echo $rest | sed 's/ / | /g')"
done | sort
