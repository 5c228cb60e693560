#!/bin/bash
# SASS of the shipped closest-hit kernel (k_trace) out of voidray_b200/libvoidray_cuda.so -> profiles/<tag>_ktrace_sass.txt:
# an opcode histogram first, then the listing without the encoding words.
TAG=${1:-r2}
OUT=profiles/${TAG}_ktrace_sass.txt
TMP=$(mktemp)
cuobjdump -sass voidray_b200/libvoidray_cuda.so 2>/dev/null | awk '/Function : .*7k_traceENS/{f=1} f&&/Function : /&&!/7k_traceENS/{f=0} f' \
  | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's|/\* 0x[0-9a-f]* \*/||; s/[[:space:]]*$//; s/^\s*//' > $TMP
{
  echo "# k_trace (closest hit) as shipped: $(wc -l < $TMP) SASS instructions, $(git rev-parse --short HEAD 2>/dev/null) + working tree"
  echo "# nvcc $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo"
  echo "# opcode histogram:"
  awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); n[op]++} END {for (k in n) printf "#   %-10s %d\n", k, n[k]}' $TMP | sort -k3 -n -r
  echo "#"
  cat $TMP
} > $OUT
rm -f $TMP
head -5 $OUT; wc -l $OUT
