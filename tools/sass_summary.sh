#!/bin/bash
# per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): tcgen05.mma = UTC*MMA,
# tcgen05.ld / st = LDTM / STTM, tcgen05.commit = UTCBAR, TMA = UTMALDG / UTMASTG / UBLKCP (bulk copy) / UBLKPF (bulk L2 prefetch)
LIB=${1:-pde_inverse_problem_b200/libpdeip.so}
echo "# cuobjdump -sass $LIB  ($(date -u +%F), nvcc $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# kernel | UTCHMMA | LDTM | STTM | UTCBAR | UTMALDG | UTMASTG | UBLKCP | UBLKPF | HMMA | FFMA2 | MUFU | total instructions"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { if (name != "") print_row(); name = $3; for (k in c) delete c[k]; tot = 0; next }
  /^ +\/\*[0-9a-f][0-9a-f]*\*\/ +[A-Z@]/ { tot++
    if ($0 ~ /UTCHMMA/) c["UTCHMMA"]++; if ($0 ~ /LDTM/) c["LDTM"]++; if ($0 ~ /STTM/) c["STTM"]++
    if ($0 ~ /UTCBAR/) c["UTCBAR"]++; if ($0 ~ /UTMALDG/) c["UTMALDG"]++; if ($0 ~ /UTMASTG/) c["UTMASTG"]++
    if ($0 ~ /UBLKCP/) c["UBLKCP"]++; if ($0 ~ /UBLKPF/) c["UBLKPF"]++; if ($0 ~ / HMMA/) c["HMMA"]++
    if ($0 ~ /FFMA2/) c["FFMA2"]++; if ($0 ~ /MUFU/) c["MUFU"]++ }
  function print_row() { printf "%s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d\n", name, c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTCBAR"], c["UTMALDG"], c["UTMASTG"], c["UBLKCP"], c["UBLKPF"], c["HMMA"], c["FFMA2"], c["MUFU"], tot }
  END { if (name != "") print_row() }' | while IFS= read -r line; do
    n=$(echo "$line" | cut -d'|' -f1 | tr -d ' '); rest=$(echo "$line" | cut -d'|' -f2-)
    echo "$(echo "$n" | c++filt | sed 's/(.*//') |$rest"
  done | sort
