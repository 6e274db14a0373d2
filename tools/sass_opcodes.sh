#!/bin/bash
# Evidence that the shipped kernels are Blackwell-native: counts of the SASS mnemonics that correspond to tcgen05 / TMEM / TMA
# (B200_PROFILING.md "What proves a Blackwell-native kernel") per kernel family of libggufb200.so.
#   tcgen05.mma -> UTCHMMA   tcgen05.ld/st -> LDTM/STTM   cp.async.bulk.tensor load / store -> UTMALDG / UTMASTG   cp.async.bulk -> UBLKCP
#   tcgen05.commit -> UTCBAR   mma.sync -> HMMA (legacy tensor path, the M <= 8 exact GEMV only)
cd "$(dirname "$0")/../comfyui-gguf_b200/csrc" || exit 1
SO=libggufb200.so
echo "# cuobjdump -sass $SO ($(stat -c %s $SO) bytes, built $(date -u -r $SO +%Y-%m-%dT%H:%MZ)); instruction counts summed over all template instantiations of a kernel"
printf "%-22s %8s %9s %7s %7s %8s %8s %8s %8s %7s %9s\n" kernel instances UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKCP UTCBAR HMMA total_instr
cuobjdump -sass $SO | awk '
/Function : / { name=$3; fam="other";
  if (name ~ /dequant_kernel/) fam="dequant_kernel"; else if (name ~ /unpack_kernel/) fam="unpack_kernel";
  else if (name ~ /rows_kernel/) fam="rows_kernel"; else if (name ~ /gemv_mma_kernel/) fam="gemv_mma_kernel"; else if (name ~ /gemv2_kernel/) fam="gemv2_kernel";
  else if (name ~ /gemv_bf16w/) fam="gemv_bf16w_kernel"; else if (name ~ /gemm2_kernel/) fam="gemm2_kernel";
  else if (name ~ /gemm3_kernel/) fam="gemm3_kernel"; else if (name ~ /gemm4_kernel/) fam="gemm4_kernel";
  else if (name ~ /finalize/) fam="finalize_kernels"; else if (name ~ /repack_kernel/) fam="repack_kernel";
  inst[fam]++; next }
/^[ \t]+\/\*[0-9a-f]{4}\*\// { tot[fam]++;
  if ($0 ~ /UTCHMMA/) a[fam]++; if ($0 ~ /LDTM/) b[fam]++; if ($0 ~ /STTM/) c[fam]++; if ($0 ~ /UTMALDG/) d[fam]++; if ($0 ~ /UTMASTG/) h[fam]++;
  if ($0 ~ /UBLKCP/) e[fam]++; if ($0 ~ /UTCBAR/) f[fam]++; if ($0 ~ / HMMA/) g[fam]++ }
END { for (k in inst) printf "%-22s %8d %9d %7d %7d %8d %8d %8d %8d %7d %9d\n", k, inst[k], a[k], b[k], c[k], d[k], h[k], e[k], f[k], g[k], tot[k] }' | sort
