#!/bin/bash
# Evidence that survives the round: SASS mnemonic counts of the shipped objects -> profiles/r2_sass_counts.txt
cd "$(dirname "$0")/../miniframe_b200/csrc" || exit 1
out=../../profiles/r2_sass_counts.txt
{
  echo "# cuobjdump -sass | grep -c, objects built by miniframe_b200/csrc/Makefile (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), sm_100a)"
  echo "# source sha256: $(sha256sum mf_potrf.cu mf_cov.cu | awk '{print substr($1,1,16), $2}' | tr '\n' ' ')"
  for obj in mf_potrf.o mf_cov.o mf_solve.o mf_grad.o; do
    [ -f $obj ] || continue
    cuobjdump -sass $obj > /tmp/_sass.txt
    echo "== $obj"
    for m in DMMA UTMALDG UTMASTG UBLKPF 'STG.E.128' 'STG.E.64' 'LDG.E.128' 'LDS.128' SYNCS 'BAR.SYNC' 'MUFU.RSQ64H' DFMA SHFL LDL STL; do
      printf "  %-12s %6d\n" "$m" "$(grep -c -- "$m" /tmp/_sass.txt)"
    done
    echo "  per kernel (DMMA / UTMALDG / UTMASTG / BAR.SYNC):"
    awk '/Function :/ {name=$3} /DMMA/ {d[name]++} /UTMALDG/ {l[name]++} /UTMASTG/ {s[name]++} /BAR.SYNC/ {b[name]++} END {for (k in d) printf "    %s %d %d %d %d\n", k, d[k], l[k]+0, s[k]+0, b[k]+0; for (k in s) if (!(k in d)) printf "    %s 0 %d %d %d\n", k, l[k]+0, s[k], b[k]+0}' /tmp/_sass.txt | c++filt | cut -c1-160
  done
} > $out
cat $out
