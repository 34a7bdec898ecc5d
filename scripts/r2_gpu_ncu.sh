#!/bin/bash
# Round 2: `ncu --set full` captures of the kernels the north_star names beside the stencil — scatter / sort (assembly), the
# data-term and update kernels, the 2D TMA kernel — summarised into gpurun_out/*.md (copied to profiles/ afterwards).
set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:scatter_points|radix_scatter|radix_hist|cell_keys|canonicalise|diagonal_kernel|apply_blocks_kernel|pcg_update_kernel|stencil3d_tma" -c 60 \
    -o gpurun_out/r2i_asm python scripts/profile_step.py 512 6 > gpurun_out/r2i_asm_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r2i_asm.ncu-rep > gpurun_out/r2i_full_512_f32_assembly_and_iteration.md 2>&1; grep -c "^###" gpurun_out/r2i_full_512_f32_assembly_and_iteration.md
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:stencil2d|pcg_update_kernel|apply_blocks" -c 24 \
    -o gpurun_out/r2i_2d python scripts/profile_2d.py > gpurun_out/r2i_2d_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r2i_2d.ncu-rep > gpurun_out/r2i_full_2048sq_f32_2d_kernel.md 2>&1; grep -c "^###" gpurun_out/r2i_full_2048sq_f32_2d_kernel.md
rm -f gpurun_out/r2i_asm.ncu-rep gpurun_out/r2i_2d.ncu-rep
