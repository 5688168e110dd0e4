#!/bin/bash
# second part of the small-kernel capture: input kernels (projection, resize, voxelizer), readers, devoxelization, sampling, class
# embeddings / tokens, fused decoder
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
LS3D_PROFILE_INPUTS=1 timeout 400 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
    --profile-from-start off \
    -k regex:'vox_|project_points|resize_u8|pad3|vfe_|three_nn|three_interpolate|sample_image|ce_partial|ce_max|ce_final|class_tokens|sffm_decoder' \
    -c 60 --csv --page raw --log-file $O/prof_small_kernels2_raw.csv python scripts/prof_unet_step.py > $O/ncu_small2.log 2>&1; tail -n 2 $O/ncu_small2.log
wc -l $O/prof_small_kernels2_raw.csv
