#!/bin/sh
# Counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md "What proves a Blackwell-native kernel"),
# per object file of libwhisper_b200.so.  Usage: tools/sass_counts.sh > profiles/rNN_sass_counts.txt   (after make -C whisper_apr_b200/csrc)
cd "$(dirname "$0")/.." || exit 1
echo "# cuobjdump -sass build/obj/<file>.o | grep -c <mnemonic>   ($(nvcc --version | tail -1); operand format: see wb_operand_format())"
printf "%-18s %8s %8s %8s %9s %6s %6s %7s %6s %6s\n" object UTCHMMA UTMALDG UTMASTG UTMAREDG LDTM STTM UTCBAR legacyHMMA MUFU.EX2
for f in gemm attention mel elementwise decoder audio_pre loader pipeline; do
  o=build/obj/$f.cu.o
  [ -f "$o" ] || continue
  s=$(cuobjdump -sass "$o")
  c() { printf "%s" "$s" | grep -c "$1"; }
  printf "%-18s %8s %8s %8s %9s %6s %6s %7s %6s %6s\n" "$f.cu" "$(c UTCHMMA)" "$(c UTMALDG)" "$(c UTMASTG)" "$(c UTMAREDG)" "$(c LDTM)" "$(c STTM)" "$(c UTCBAR)" "$(c '[^C]HMMA')" "$(c 'MUFU.EX2')"
done
echo
echo "# variants inside gemm.cu.o (UTCHMMA.2CTA = tcgen05.mma cta_group::2; UTMALDG.3D.2CTA = TMA loads crediting the leader CTA's barrier)"
cuobjdump -sass build/obj/gemm.cu.o | grep -oE "UTCHMMA[.A-Z0-9]*|UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|UTMAREDG[.A-Z0-9]*|UTCBAR[.A-Z0-9]*|LDTM[.a-zA-Z0-9]*" | sort | uniq -c
echo
echo "# variants inside attention.cu.o (STTM = tcgen05.st of P into tensor memory; UTCHMMA with a TMEM A operand)"
cuobjdump -sass build/obj/attention.cu.o | grep -oE "UTCHMMA[.A-Z0-9]*|UTMALDG[.A-Z0-9]*|LDTM[.a-zA-Z0-9]*|STTM[.a-zA-Z0-9]*|UTCBAR[.A-Z0-9]*" | sort | uniq -c
echo
echo "# packed fp32x2 arithmetic in mel.cu.o (one instruction per complex operation; scalar counterparts beside them)"
cuobjdump -sass build/obj/mel.cu.o | grep -oE "\b(FADD2|FMUL2|FFMA2|FADD|FMUL|FFMA)\b" | sort | uniq -c
