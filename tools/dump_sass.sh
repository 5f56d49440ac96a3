#!/bin/bash
# cuobjdump -sass of the kernels the roofline / VERDICT name, from the in-tree objects
# (amg-ann_b200/build/*.o, sm_100a) into profiles/<round>_sass/.  No GPU needed.
set -eu
R=${1:-r2}
cd "$(dirname "$0")/../amg-ann_b200"
OUT=../profiles/${R}_sass
mkdir -p $OUT
dump() { cuobjdump -sass -fun "$2" "$1" | grep -v '^\s*$' > "$OUT/$3.sass"; echo "$3: $(grep -c ';' $OUT/$3.sass) instructions"; }
dump build/amgb_solve.o _ZN4amgb16sell_rows_kernelILi1ELb1ENS_9EpiJacobiEEEviiiiPKiS3_PKdS5_S5_iT1_ sell_rows_kernel_T1_SPLIT_EpiJacobi
dump build/amgb_solve.o _ZN4amgb16sell_rows_kernelILi1ELb0ENS_11EpiResidualEEEviiiiPKiS3_PKdS5_S5_iT1_ sell_rows_kernel_T1_EpiResidual
dump build/amgb_solve.o _ZN4amgb20sell_spmv_dot_kernelILi1EEEviiiiPKiS2_PKdS4_PdS5_ sell_spmv_dot_kernel_T1
dump build/amgb_setup.o _ZN4amgb21spgemm_numeric_kernelILi8ELi128ELb0EEEvlPKiS2_PKdS2_S2_S4_S2_PiPdS5_S5_ spgemm_numeric_kernel_G8_CAP128
dump build/amgb_setup.o _ZN4amgb21spgemm_numeric_kernelILi32ELi512ELb1EEEvlPKiS2_PKdS2_S2_S4_S2_PiPdS5_S5_ spgemm_numeric_kernel_G32_CAP512_SORT
dump build/amgb_setup.o _ZN4amgb24interp_fill_group_kernelElPKiS1_PKdPKhS1_S1_S1_S1_S3_S1_PiPdlPhS6_ interp_fill_group_kernel
dump build/amgb_setup.o _ZN4amgb18interp_fill_kernelElPKiS1_PKdPKhS1_S1_S3_S1_PiPdlS5_ interp_fill_kernel
dump build/amgb_pool.o _ZN4amgb19pool_entries_kernelILb0EEEvNS_6BinMapEixiPKiS3_PKdPdPxS6_S6_ pool_entries_kernel
dump build/amgb_tail.o _ZN4amgb17cycle_tail_kernelENS_8TailDescEPKhPKdS4_Pd cycle_tail_kernel
dump build/amgb_setup.o _ZN4amgb18spgemm_flat_kernelILi8ELi256ELi256ELb0ELb0ELb0EEEvlPKiS2_PKdS2_S2_S4_S2_PiPdS5_S5_S5_ spgemm_flat_kernel_G8_count
dump build/amgb_setup.o _ZN4amgb18spgemm_flat_kernelILi8ELi128ELi256ELb1ELb0ELb0EEEvlPKiS2_PKdS2_S2_S4_S2_PiPdS5_S5_S5_ spgemm_flat_kernel_G8_numeric_flattened
dump build/amgb_setup.o _ZN4amgb18spgemm_flat_kernelILi8ELi128ELi256ELb1ELb0ELb1EEEvlPKiS2_PKdS2_S2_S4_S2_PiPdS5_S5_S5_ spgemm_flat_kernel_G8_numeric_entrywise
