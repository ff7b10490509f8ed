// oracle/ref_kernels_gpu.cpp -- TEST INFRASTRUCTURE.
// Compiles the reference's own "GPU program" (/root/reference/src/GPU_kernels.cl) for the
// host CPU through clc_compat.hpp.  GPU_kernels.inc is produced in a temporary directory
// by oracle/Makefile (sed rewrites only the OpenCL vector-literal cast "(int4)(a,b,c,d)" to
// the constructor call "int4(a,b,c,d)"; nothing else is touched) and is never stored in
// the repository.
#include "clc_compat.hpp"

namespace clc {
namespace gpu_prog {
#include "GPU_kernels.inc"
}  // namespace gpu_prog
}  // namespace clc

extern "C" const clc::kernel_desc vp8ref_gpu_kernels[] = {
    CLC_KERNEL_ENTRY(clc::gpu_prog, reset_vectors),
    CLC_KERNEL_ENTRY(clc::gpu_prog, downsample_x2),
    CLC_KERNEL_ENTRY(clc::gpu_prog, luma_search_1step),
    CLC_KERNEL_ENTRY(clc::gpu_prog, luma_search_2step),
    CLC_KERNEL_ENTRY(clc::gpu_prog, select_reference),
    CLC_KERNEL_ENTRY(clc::gpu_prog, prepare_predictors_and_residual),
    CLC_KERNEL_ENTRY(clc::gpu_prog, pack_8x8_into_16x16),
    CLC_KERNEL_ENTRY(clc::gpu_prog, dct4x4),
    CLC_KERNEL_ENTRY(clc::gpu_prog, wht4x4_iwht4x4),
    CLC_KERNEL_ENTRY(clc::gpu_prog, idct4x4),
    CLC_KERNEL_ENTRY(clc::gpu_prog, count_SSIM_luma),
    CLC_KERNEL_ENTRY(clc::gpu_prog, count_SSIM_chroma),
    CLC_KERNEL_ENTRY(clc::gpu_prog, gather_SSIM),
    {nullptr, 0, nullptr, nullptr}};

// direct access to the reference's block-cost helper, for the Q1 unit test
extern "C" int vp8ref_weight_opt(const int *residual16) {
    clc::int16 L;
    clc::int4 XX;
    for (int i = 0; i < 16; ++i) L.s[i] = residual16[i];
    clc::gpu_prog::weight_opt(&XX, &L);
    return XX.x;
}
