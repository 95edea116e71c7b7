#!/bin/sh
# Counts the SASS mnemonics that prove which hardware paths the built objects use
# (B200_PROFILING.md "What proves a Blackwell-native kernel").  Run after build():
#   sh tools/sass_evidence.sh > profiles/r01b_sass_evidence.txt
cd "$(dirname "$0")/../cqs_b200/csrc" || exit 1
PAT='\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|SYNCS\.[A-Z_]+|MATCH\.ANY|STG\.E\.STRONG\.SYS|LDG\.E\.STRONG\.SYS|MEMBAR\.[A-Z]+\.SYS|HMMA|HGMMA)\b'
for f in scan_batch.o scan_v_m0_small.o scan_v_m1_large.o peer.o sparse_fuse.o sparse_build.o; do
  echo "== $f"
  /usr/local/cuda/bin/cuobjdump -sass "$f" 2>/dev/null | grep -oE "$PAT" | sort | uniq -c | sort -rn
done
