#!/bin/bash
# SASS evidence that the kernels are Blackwell-native (B200_PROFILING.md "What proves a Blackwell-native kernel"):
# per kernel, the count of tcgen05 / TMEM / TMA / legacy-MMA instructions in the in-tree library.
#   tools/sass_evidence.sh > profiles/rNN_sass_evidence.txt
LIB=${1:-hypernet_image_captioning_b200/lib/libcaphn_b200.so}
echo "# cuobjdump -sass $LIB  (sm_100a) -- instruction counts per kernel"
echo "# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store, UBLKCP = cp.async.bulk,"
echo "# UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, SYNCS = mbarrier ops"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; next }
{ for (i=1;i<=NF;i++) { t=$i; sub(/\..*/,"",t);
    if (t=="UTCHMMA"||t=="LDTM"||t=="UTMALDG"||t=="UTMASTG"||t=="UBLKCP"||t=="UTCBAR"||t=="HMMA"||t=="SYNCS"||t=="UTCQMMA"||t=="STTM") c[fn" "t]++ } }
END { for (k in c) print k, c[k] }' | sort | while read fn ins n; do
  printf "%-110s %-8s %5d\n" "$(echo $fn | c++filt | cut -c1-110)" "$ins" "$n"; done
echo
echo "# excerpt: first tcgen05 / TMA instructions of gemm_tc_kernel<true,false>"
cuobjdump -sass "$LIB" | awk '/Function :.*gemm_tc_kernelILb1ELb0/ {p=1} p && /UTCHMMA|LDTM|UTMALDG|UTMASTG|UTCBAR/ {print; n++} n>=24 {exit}'
