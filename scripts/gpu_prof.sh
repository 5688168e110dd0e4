#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
rm -f $O/*.ncu-rep
P=${1:-1}
LS3D_PRECISE=$P timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gather_gemm \
    -o $O/prof_gemm_p$P -f python scripts/prof_gemm.py > $O/ncu_gemm_p$P.log 2>&1
ncu -i $O/prof_gemm_p$P.ncu-rep --page raw --csv > $O/prof_gemm_p${P}_raw.csv 2>/dev/null
ncu -i $O/prof_gemm_p$P.ncu-rep --page source --csv > $O/prof_gemm_p${P}_source.csv 2>/dev/null
ncu -i $O/prof_gemm_p$P.ncu-rep --page details > $O/prof_gemm_p${P}_details.txt 2>/dev/null
ls -la $O/*.ncu-rep
SZ=$(stat -c %s $O/prof_gemm_p$P.ncu-rep); if [ "$SZ" -gt 40000000 ]; then rm -f $O/prof_gemm_p$P.ncu-rep; fi
timeout 300 python scripts/diag_gemm.py > $O/diag_gemm.log 2>&1
tail -n 3 $O/ncu_gemm_p$P.log; grep timing $O/diag_gemm.log | head -8
