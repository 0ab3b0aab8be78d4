#!/bin/bash
# profiles/sass_grep.txt: which sm_100a instructions the hot kernels really contain (cuobjdump -sass of the built objects).
L=single-shot-detector_b200/lib/obj
OUTF=${1:-profiles/sass_grep.txt}
count() {  # object, function-substring, mnemonic regex
  cuobjdump -sass $L/$1.o | awk -v f="$2" '/Function :/ {on = index($0, f) > 0} on' | grep -cE "$3"
}
{
echo "SASS evidence (cuobjdump -sass of $L/*.o, built by csrc/Makefile with -gencode arch=compute_100a,code=sm_100a): count of"
echo "instructions per kernel.  No tensor-core instruction is expected anywhere (nothing on this path is a contraction)."
echo
cuobjdump -sass $L/loss.o | grep -m1 "arch ="
echo
printf "%-58s %8s %8s %8s %8s %8s %8s %8s %8s\n" "kernel" "LDG.128" "UBLKCP" "SYNCS" "FFMA2" "FMUL2" "MUFU" "REDUX" "ATOM/RED"
row() {
  printf "%-58s %8s %8s %8s %8s %8s %8s %8s %8s\n" "$3" "$(count $1 $2 'LDG\.E(\.[A-Z]+)*\.128')" "$(count $1 $2 'UBLKCP')" "$(count $1 $2 'SYNCS')" \
     "$(count $1 $2 'FFMA2')" "$(count $1 $2 'FMUL2')" "$(count $1 $2 'MUFU')" "$(count $1 $2 'REDUX')" "$(count $1 $2 ' ATOM|ATOMG|ATOMS| RED\.')"
}
row train_step _Z17train_step_kernelILi0E "train_step_kernel<gamma=2> (train_step.cu)"
row head _Z16head_flat_kernelILi0ELb0E "head_flat_kernel<gamma=2, forward> (head.cu)"
row head _Z16head_flat_kernelILi0ELb1E "head_flat_kernel<gamma=2, forward+backward> (head.cu)"
row head _Z16head_rows_kernelILi0ELb0E "head_rows_kernel<gamma=2, forward> (head.cu)"
row loss _Z15ssd_loss_kernelILi0ELb0E "ssd_loss_kernel<gamma=2> (loss.cu, TMA ring)"
row loss_backward _Z24ssd_loss_backward_kernelILi0ELb1E "ssd_loss_backward_kernel<gamma=2, with loss> (loss_backward.cu)"
row matcher match_kernel "match_kernel (all variants, matcher.cu)"
row postprocess _Z13filter_kernelILb1E "filter_kernel<logits> (postprocess.cu)"
row postprocess _Z18head_filter_kernelILb1E "head_filter_kernel<logits> (postprocess.cu)"
row postprocess _Z19filter_dense_kernelILb1E "filter_dense_kernel<logits> (postprocess.cu)"
row postprocess _Z16nms_small_kernelILb0E "nms_small_kernel (postprocess.cu)"
row postprocess _Z10nms_kernelILb0E "nms_kernel (postprocess.cu)"
row postprocess _Z17nms_rounds_kernelILb0ELb1E "nms_rounds_kernel<logits> (postprocess.cu)"
row postprocess pack_kernel "pack_kernel + pack_by_label_kernel (postprocess.cu)"
row comm comm "comm kernels (comm.cu)"
echo
echo "tensor-core / TMEM mnemonics over the whole library (UTCMMA, UTCHMMA, HMMA, IMMA, tcgen05): $(for f in $L/*.o; do cuobjdump -sass $f; done | grep -cE 'UTC[A-Z]*MMA|HMMA|IMMA|UTCBAR|LDTM|STTM')"
echo
echo "--- excerpt: steady-state loop of filter_kernel<logits> (six 128-bit no-allocate loads, one max tree, one branch) ---"
cuobjdump -sass $L/postprocess.o | awk '/Function :/ {on = index($0, "_Z13filter_kernelILb1E") > 0} on' | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##' | \
  awk '/LDG.E.NA.128/ && !s {s=1} s && n<34 {print; n++}' | cut -c1-100
echo
echo "--- excerpt: full-chunk loop of train_step_kernel<gamma=2> (four 128-bit loads, FMUL2 / MUFU.EX2 / FFMA2 Horner chains) ---"
cuobjdump -sass $L/train_step.o | awk '/Function :/ {on = index($0, "_Z17train_step_kernelILi0E") > 0} on' | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##' | \
  awk '/LDG.E.NA.128/ && !s {s=1} s && n<48 {print; n++}' | cut -c1-100
echo
echo "--- excerpt: producer warp of ssd_loss_kernel<gamma=2> (TMA bulk copies + mbarrier) ---"
cuobjdump -sass $L/loss.o | awk '/Function :/ {on = index($0, "_Z15ssd_loss_kernelILi0ELb0E") > 0} on' | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##' | grep -E "UBLKCP|SYNCS" | head -12 | cut -c1-110
} > $OUTF
wc -l $OUTF
