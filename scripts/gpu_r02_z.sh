#!/bin/bash
# ncu (SpeedOfLight + memory workload + launch + occupancy sections) of every NON-GEMM, NON-conv kernel of one eager MSeg3D forward:
# achieved DRAM GB/s / % of peak per kernel (north-star: each kernel evidenced by a committed capture)
cd "$(dirname "$0")/.."
python -c "import torch"
O=gpurun_out; mkdir -p $O
timeout 540 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none \
    --profile-from-start off \
    -k regex:'vox_|vfe_|grid_|nbr_|scan_|three_nn|three_interpolate|sample_image|ce_|class_tokens|upsample_sum|project_points|resize_u8|normalize|pad3|tile_plan|sffm_decoder|fill_|cast_' \
    -c 120 --csv --page raw --log-file $O/prof_small_kernels_raw.csv python scripts/prof_unet_step.py > $O/ncu_small.log 2>&1; tail -n 2 $O/ncu_small.log
wc -l $O/prof_small_kernels_raw.csv
