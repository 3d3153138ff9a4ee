#!/bin/bash
# CLI end to end on a generated paired FASTQ set: wall time of abyss-bloom-dbg with the background parser
# (ABB_STREAM_STATS shows how long the GPU thread waited for the host parser).
set -e
N=${1:-2000000}
D=${2:-/tmp/cli_probe}
mkdir -p $D
python - <<PY
from abyss_b200.synth import ReadSet
import time
t=time.time()
rs = ReadSet.from_coverage(5, int($N*150/40), 40, 150, 0.005)
h = rs.n // 2
rs.write_fastq("$D/r1.fq", 0, h); rs.write_fastq("$D/r2.fq", h, rs.n)
print("generated", rs.n, "reads in %.1fs" % (time.time()-t))
PY
ls -la $D/*.fq
for j in 1 8; do
  s=$(date +%s.%N)
  ABB_STREAM_STATS=1 abyss_b200/lib/abyss-bloom-dbg -k64 --kc=3 -b1G -H4 -j$j $D/r1.fq $D/r2.fq > $D/out_j$j.fa 2> $D/err_j$j.txt
  e=$(date +%s.%N)
  echo "-j$j: wall $(echo "$e - $s" | bc -l 2>/dev/null || python -c "print($e-$s)") s, unitigs $(grep -c '>' $D/out_j$j.fa)"; grep BatchStream $D/err_j$j.txt
done
cmp $D/out_j1.fa $D/out_j8.fa && echo "same output"
