// oracle/ref_kernels_cpu.cpp -- TEST INFRASTRUCTURE.
// Compiles the reference's own "CPU program" (/root/reference/src/CPU_kernels.cl, built with
// -DLOOP_FILTER as src/init.h:337 does in the default mode) for the host CPU through
// clc_compat.hpp.  See ref_kernels_gpu.cpp for how CPU_kernels.inc is produced.
#define LOOP_FILTER 1
#include "clc_compat.hpp"

namespace clc {
namespace cpu_prog {
#include "CPU_kernels.inc"
}  // namespace cpu_prog
}  // namespace clc

extern "C" const clc::kernel_desc vp8ref_cpu_kernels[] = {
    CLC_KERNEL_ENTRY(clc::cpu_prog, encode_coefficients),
    CLC_KERNEL_ENTRY(clc::cpu_prog, count_probs),
    CLC_KERNEL_ENTRY(clc::cpu_prog, num_div_denom),
    CLC_KERNEL_ENTRY(clc::cpu_prog, prepare_filter_mask),
    CLC_KERNEL_ENTRY(clc::cpu_prog, loop_filter_frame_luma),
    CLC_KERNEL_ENTRY(clc::cpu_prog, loop_filter_frame_chroma),
    {nullptr, 0, nullptr, nullptr}};
