#!/bin/sh
# Census of Blackwell-only (tcgen05 / TMEM / TMA / mbarrier) SASS in the shipped library, per kernel.
#   sh profiles/scripts/sass_census.sh > profiles/r02_sass_blackwell_census.txt
LIB=${1:-multimodalworddiscovery_b200/libmwd_b200.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%d)), nvcc $(nvcc --version | grep -o 'release [0-9.]*')"
echo "# UTCHMMA = tcgen05.mma (kind::tf32/f16), UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk (bulk store),"
echo "# LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier, REDG.F64 = red.global.add.f64"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3 }
{
  for (i = 1; i <= NF; i++) {
    op=$i
    if (op ~ /^(UTCHMMA|UTCMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS|REDG|DMMA|DFMA|FFMA|UTCCP)/) {
      split(op, p, "."); key=p[1]; if (key=="REDG" || key=="SYNCS" || key=="LDTM" || key=="UTMALDG" || key=="UBLKCP") key=op
      sub(/;$/, "", key); cnt[fn SUBSEP key]++; keys[key]=1; fns[fn]=1
    }
  }
}
END {
  for (f in fns) {
    line=""
    tc=0
    for (k in keys) if ((f SUBSEP k) in cnt) { line=line sprintf(" %s=%d", k, cnt[f SUBSEP k]); if (k ~ /^(UTC|UTMA|UBLKCP|LDTM|SYNCS)/) tc=1 }
    if (tc) print f ":" line
  }
}' | sort | c++filt | sed 's/mwd::(anonymous namespace):://' | cut -c1-400
echo "# totals over the whole library"
cuobjdump -sass "$LIB" | grep -o "UTCHMMA\|UTMALDG[.0-9A-Z]*\|UBLKCP[.A-Z]*\|LDTM[.a-z0-9]*\|UTCBAR\|DMMA\|REDG.E.ADD.F64[.A-Z]*" | sort | uniq -c
