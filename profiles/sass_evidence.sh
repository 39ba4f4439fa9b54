#!/bin/bash
# SASS evidence for the default fused kernels of the built library: TMA (UTMALDG = cp.async.bulk.tensor,
# UBLKCP = cp.async.bulk), mbarrier (SYNCS.*), warp reductions (CREDUX / REDUX), async-proxy fences.
#   bash profiles/sass_evidence.sh > profiles/r2_sass_excerpt.txt
so=${1:-peakachu_b200/libpeakachu_b200.so}
for k in ILi5ELi256ELi2ELi4224ELi4ELi1ELi512ELi4ELi0ELi0ELi0E ILi7ELi112ELi2ELi6400ELi4ELi1ELi384ELi12ELi0ELi1ELi2E; do
  echo "== k_score_fused$k  ($(cuobjdump -sass $so | awk -v k="$k" '/Function :/ {on = index($0, "k_score_fused" k) > 0} on' | grep -c '^ *\/\*[0-9a-f]*\*\/') SASS instructions, target $(cuobjdump -lelf $so | grep -o 'sm_[0-9a-z]*' | sort -u | tr '\n' ' '))"
  cuobjdump -sass $so | awk -v k="$k" '/Function :/ {on = index($0, "k_score_fused" k) > 0} on' \
    | grep -E "UTMALDG|UBLKCP|SYNCS\.|CREDUX|REDUX|FENCE\.VIEW\.ASYNC|BAR\.SYNC|DFMA|LDS\.64|DMUL" \
    | sed -E 's/^ *//; s/ *\/\* 0x[0-9a-f]* \*\/$//' \
    | awk '{op=$2; if ($2 ~ /^@/) op=$3; sub(/\..*/, "", op); n[op]++; if (shown[op]++ < 3) print "   " $0} END {printf "   counts:"; for (o in n) printf " %s=%d", o, n[o]; print ""}'
done
