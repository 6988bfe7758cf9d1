#!/bin/bash
# SASS evidence of the sm_100a-specific instructions in libpctrans_msda.so (authoring container):
#   bash scripts/sass_excerpt.sh > profiles/r02_sass_excerpt.txt
LIB=${1:-pctrans_b200/_build/libpctrans_msda.so}
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo "# cuobjdump -sass $LIB  ($(grep -c 'Function :' "$TMP") kernels, $(wc -l < "$TMP") lines)"
echo "# arch: $(grep -m1 'arch =' "$TMP")"
echo
echo "## mnemonic counts over the whole library"
for m in UTCHMMA 'UTCHMMA.2CTA' LDTM STTM UBLKCP UTCBAR SYNCS 'REDG.E.ADD.F32x4' 'REDG.E.ADD.64' 'REDG.E.ADD.BF16x2' FFMA2 FMUL2 'LDG.E.128' 'LDG.E.ENL2.256' 'LDGSTS' 'ATOMS.ADD' 'ATOMS.POPC.INC' 'ATOMS.CAST.SPIN' CREDUX SHFL.BFLY SHFL.IDX 'CCTL' ; do
  printf "%-22s %6d\n" "$m" "$(grep -c -- "$m" "$TMP")"
done
echo
excerpt() {  # $1 = kernel-name regex, $2 = mnemonic regex, $3 = max lines
  awk -v k="$1" '/Function :/{f=($0 ~ k)} f' "$TMP" | grep -E -- "$2" | head -"$3" | sed 's/^\s*//'
}
echo "## msda_bwd_binned_kernel<float,8>: shared-memory integer atomics (counting sort), cp.async staging, packed fp32 math, vector reductions"
excerpt 'msda_bwd_binned_kernelIfLi8E' 'ATOMS|LDGSTS|FFMA2|FMUL2|REDG|CREDUX' 24
echo
echo "## msda_bwd_vec_kernel<float,8,16>: 128-bit vector reductions and loads"
excerpt 'msda_bwd_vec_kernelIfLi8ELi16ELb0Ef' 'REDG|LDG.E.128' 10
echo
echo "## msda_fwd_vec_kernel<float,8,16>: 128-bit gathers, shared-memory tap exchange, streaming store"
excerpt 'msda_fwd_vec_kernelIfLi8ELi16ELb0' 'LDG.E.128|LDS.128|STS.128|STG' 12
echo
echo "## linear_tf32x3_kernel: tcgen05 MMA (UTCHMMA), TMEM loads/stores (LDTM/STTM), bulk copies (UBLKCP)"
excerpt 'linear_tf32x3_kernel' 'UTCHMMA|LDTM|STTM|UBLKCP|UTCBAR|SYNCS' 20
echo
echo "## linear_bf16_kernel / wgrad_bf16_kernel: the same MMAs with bf16 operands (kind::f16), bf16x2 packing of the loaded activations"
excerpt 'linear_bf16_kernel' 'UTCHMMA|F2FP|UBLKCP' 10
excerpt 'wgrad_bf16_kernelILb0ELb0' 'UTCHMMA|F2FP' 6
echo
echo "## linear_tf32x3_pair_kernel: cta_group::2 MMAs"
excerpt 'linear_tf32x3_pair_kernel' 'UTCHMMA' 6
rm -f "$TMP"
