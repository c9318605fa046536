#!/bin/bash
# SASS / ptxas evidence for profiles/ (runs in the dev container: no GPU needed)
SO=lyricalignment_b200/_C/liblyricalign.so
echo "# SASS mnemonics per kernel of $SO (cuobjdump -sass; CUDA $(nvcc --version | grep -o 'release [0-9.]*'))"
echo "# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA),"
echo "# SYNCS = mbarrier, ELECT = elect.sync, FFMA2/FADD2 = packed fp32x2, DADD/DSETP = the fp64 DP, MUFU.EX2/LG2 = SFU"
cuobjdump -sass $SO 2>/dev/null | awk '
  /Function : /{name=$3}
  { for (m in pats) if (index($0, m)) cnt[name, m]++ }
  BEGIN{ split("UTCHMMA UTCQMMA LDTM STTM UTCBAR UTCATOMSWS UBLKCP UBLKPF SYNCS ELECT FFMA2 FADD2 DADD DSETP MUFU.EX2 MUFU.LG2 SHFL LDS STS BAR.SYNC", a, " "); for (i in a) pats[a[i]]=1 }
  END{ for (k in cnt) { split(k, p, SUBSEP); print p[1], p[2], cnt[k] } }' | sort | awk '{k[$1]=k[$1] " " $2 "x" $3} END{for (n in k) print n ":" k[n]}' | sort
echo
echo "# ptxas -v (registers / spills / shared memory / barriers)"
python -m lyricalignment_b200.build --force -v 2>&1 | grep -E "Compiling entry function|Used [0-9]+ registers|bytes stack frame" | sed 's/ptxas info    : //' | paste - - - | sed "s/Compiling entry function '\(.*\)' for 'sm_100a'/\1/" 
