#!/bin/bash
# ncu --set full of (a) the fused SF-Phase decoder (source-level stall samples), (b) every conv launch of one camera forward
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sffm_decoder --launch-skip 2 -c 1 -o $O/prof_dec -f \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "sffm_decoder" > $O/ncu_dec.log 2>&1; tail -n 2 $O/ncu_dec.log
ncu -i $O/prof_dec.ncu-rep --page source --csv > $O/prof_dec_source.csv 2>/dev/null
ncu -i $O/prof_dec.ncu-rep --page details > $O/prof_dec_details.txt 2>/dev/null
LS3D_PROFILE_RANGE=1 timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:conv3x3_f16 -c 320 --csv --page raw \
    --log-file $O/prof_conv_all_raw.csv python scripts/prof_camera.py > $O/ncu_conv.log 2>&1; tail -n 3 $O/ncu_conv.log
ls -la $O | tail -8
