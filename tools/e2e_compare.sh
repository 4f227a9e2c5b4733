#!/bin/bash
# End-to-end parity run: same synthetic inputs through the reference fork (Oracle B) and the B200 build;
# compares every output line except @PG.  usage: tools/e2e_compare.sh <workdir> [n_pairs] [n_long] [genome_bp]
set -u
W=${1:-/tmp/e2e}; NP=${2:-20000}; NL=${3:-100}; G=${4:-2000000}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=$ROOT/oracle/_ref/minimap2_B; NEW=$ROOT/build/minimap2-b200; SYN=$ROOT/build/mmsynth
mkdir -p $W && cd $W
[ -f ref.fa ] || $SYN ref ref.fa $G 4 42
[ -f r1.fq ] || $SYN sr ref.fa r1.fq r2.fq $NP 44
[ -f long.fq ] || $SYN long ref.fa long.fq $NL 45
rc=0
run() { # name, args...
  local name=$1; shift
  local t0=$(date +%s.%N)
  $REF "$@" 2> ref.$name.err | grep -v '^@PG' > ref.$name.out
  local t1=$(date +%s.%N)
  $NEW "$@" 2> new.$name.err | grep -v '^@PG' > new.$name.out
  local t2=$(date +%s.%N)
  python3 -c "print('%s: ref %.2fs  new %.2fs' % ('$name', $t1 - $t0, $t2 - $t1))"
  if cmp -s ref.$name.out new.$name.out; then echo "PASS $name ($(wc -l < ref.$name.out) lines)"; else
    echo "FAIL $name: $(diff ref.$name.out new.$name.out | grep -c '^<') differing lines of $(wc -l < ref.$name.out)"; rc=1
    diff ref.$name.out new.$name.out | head -${DIFFN:-6} | cut -c1-400; grep -i "error\|assert" new.$name.err | head -5
  fi
}
T=${THREADS:-8}
run sr_paf -x sr -t $T ref.fa r1.fq r2.fq
run sr_sam -ax sr -t $T ref.fa r1.fq r2.fq
run sr_se_sam -ax sr -t $T ref.fa r1.fq
run ont_paf -x map-ont -t $T ref.fa long.fq
run ont_sam -ax map-ont -t $T ref.fa long.fq
exit $rc
