#!/bin/bash
# SASS evidence: counts of the Blackwell-specific instructions per kernel of libclica_sm100.so (run anywhere: cuobjdump)
LIB=${1:-cl-ica_b200/lib/libclica_sm100.so}
echo "| kernel | UTCHMMA (tcgen05.mma) | UTCBAR/commit | LDTM (tcgen05.ld) | UTMALDG (TMA load) | UTMASTG (TMA store) | UTMAREDG (TMA reduce) | FFMA2/FADD2/FMUL2 | MUFU.EX2 | LDGSTS (cp.async) | SYNCS (mbarrier) |"
echo "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"
cuobjdump -sass "$LIB" | awk '
/Function :/ { if (name != "") print_row(); name=$3; delete c }
function print_row() {
  n=name; gsub(/_ZN5clica[0-9]*_GLOBAL__N__[0-9a-f_]*gemm_tc_cu_[0-9a-f]*/, "", n);
  if (c["any"]>0) printf "| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n", substr(n,1,70), c["UTCHMMA"], c["UTCBAR"], c["LDTM"], c["UTMALDG"], c["UTMASTG"], c["UTMAREDG"], c["P2"], c["EX2"], c["LDGSTS"], c["SYNCS"]
}
/UTCHMMA/ {c["UTCHMMA"]++; c["any"]++} /UTCBAR/ {c["UTCBAR"]++} /LDTM/ {c["LDTM"]++} /UTMALDG/ {c["UTMALDG"]++; c["any"]++} /UTMASTG/ {c["UTMASTG"]++}
/UTMAREDG/ {c["UTMAREDG"]++} /FFMA2|FADD2|FMUL2/ {c["P2"]++; c["any"]++} /MUFU.EX2/ {c["EX2"]++} /LDGSTS/ {c["LDGSTS"]++} /SYNCS/ {c["SYNCS"]++}
END { print_row() }'
