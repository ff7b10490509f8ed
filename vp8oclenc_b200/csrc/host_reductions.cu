// The two O(N) reductions the reference host runs on every frame before it enqueues anything
// (SURVEY.md 8f-2): get_loopfilter_strength() (src/vp8enc.cpp:96-127) and the chroma differences of
// scene_change() (src/vp8enc.cpp:265-311).  They are exact integer reductions over planes that are on their way to
// the device anyway; here they are CUDA kernels behind the kernel-level C ABI.  The unmodified host cannot call them
// (both are plain C inside vp8enc.cpp, not behind an OpenCL call): INTEGRATION.md shows the three lines a maintainer
// would change.  Arithmetic follows the reference's `int` accumulators, including their wrap-around at the largest
// frame sizes (the sum of a 7680x4320 luma plane does not fit 31 bits): sums are formed modulo 2^32 and the
// divisions are C's signed ones.
#include "common.cuh"

namespace vp8 {

// acc[0] = sum of luma, acc[1] = sum over interior pixels of (p - (sum of the 8 neighbours)/8)^2,
// acc[2] = sum |last_u - cur_u|, acc[3] = sum |last_v - cur_v|        (64-bit, reduced modulo 2^32 at the end)
__global__ void __launch_bounds__(256) k_luma_statistics(const uint8_t *__restrict__ y, int width, int height,
                                                         unsigned long long *__restrict__ acc) {
    // one thread per four horizontally adjacent pixels, rows striding over the grid
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    unsigned long long sum = 0, sq = 0;
    if (x4 < width) {
        for (int row = blockIdx.y; row < height; row += gridDim.y) {
            const uint8_t *line = y + (size_t)row * width;
            const uint32_t w = *reinterpret_cast<const uint32_t *>(line + x4);
            sum += (w & 255) + ((w >> 8) & 255) + ((w >> 16) & 255) + (w >> 24);
            if (row >= 1 && row < height - 1) {
                const uint8_t *up = line - width, *dn = line + width;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = x4 + k;
                    if (x >= 1 && x < width - 1) {
                        const int nb = up[x - 1] + up[x] + up[x + 1] + line[x - 1] + line[x + 1] + dn[x - 1] + dn[x] + dn[x + 1];
                        const int d = (int)line[x] - (nb >> 3);  // (non-negative sum: /8 is the shift)
                        sq += (unsigned)(d * d);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[0], sum);
        atomicAdd(&acc[1], sq);
    }
}

__global__ void __launch_bounds__(256) k_chroma_differences(const uint8_t *__restrict__ last_u, const uint8_t *__restrict__ cur_u,
                                                            const uint8_t *__restrict__ last_v, const uint8_t *__restrict__ cur_v,
                                                            int n, unsigned long long *__restrict__ acc) {
    unsigned long long du = 0, dv = 0;
    // (n is a multiple of 4: chroma planes of a frame padded to macroblocks)
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += gridDim.x * blockDim.x * 4) {
        du += __vsadu4(*reinterpret_cast<const uint32_t *>(last_u + i), *reinterpret_cast<const uint32_t *>(cur_u + i));
        dv += __vsadu4(*reinterpret_cast<const uint32_t *>(last_v + i), *reinterpret_cast<const uint32_t *>(cur_v + i));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        du += __shfl_xor_sync(0xffffffffu, du, o);
        dv += __shfl_xor_sync(0xffffffffu, dv, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[2], du);
        atomicAdd(&acc[3], dv);
    }
}

// out[0] = reductor, out[1] = sharpness (get_loopfilter_strength), out[2] = Udiff, out[3] = Vdiff (scene_change)
__global__ void k_finish_statistics(const unsigned long long *__restrict__ acc, int width, int height, int chroma_n,
                                    int have_luma, int have_chroma, int *__restrict__ out) {
    if (have_luma) {
        const int n = width * height;
        int avg = (int)(unsigned)acc[0];       // the reference's int accumulator, modulo 2^32
        avg = (int)((unsigned)avg + (unsigned)(n / 2));
        avg /= n;
        out[0] = avg * 5 / 255 + 3;
        const int inner = (height - 1) * (width - 1);
        int div = (int)((unsigned)acc[1] + (unsigned)(inner / 2));
        div /= inner;
        int sh = div / 8;
        out[1] = sh > 7 ? 7 : sh;
    }
    if (have_chroma) {
        out[2] = (int)(unsigned)acc[2] / chroma_n;
        out[3] = (int)(unsigned)acc[3] / chroma_n;
    }
}

}  // namespace vp8

using namespace vp8;

extern "C" int vp8b200_frame_statistics(void *stream, const uint8_t *cur_y, int width, int height, const uint8_t *last_u,
                                        const uint8_t *cur_u, const uint8_t *last_v, const uint8_t *cur_v,
                                        unsigned long long *scratch4, int32_t *out4) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!scratch4 || !out4 || width < 4 || height < 3 || (width & 3)) return -(int)cudaErrorInvalidValue;
    if (cudaMemsetAsync(scratch4, 0, 4 * sizeof(unsigned long long), st) != cudaSuccess) return -(int)cudaGetLastError();
    const int chroma_n = (width / 2) * (height / 2);
    if (cur_y) {
        dim3 grid((width / 4 + 255) / 256, height < 296 ? height : 296);  // 2 x 148 row groups
        k_luma_statistics<<<grid, 256, 0, st>>>(cur_y, width, height, scratch4);
    }
    const bool chroma = last_u && cur_u && last_v && cur_v;
    if (chroma) k_chroma_differences<<<296, 256, 0, st>>>(last_u, cur_u, last_v, cur_v, chroma_n, scratch4);
    k_finish_statistics<<<1, 1, 0, st>>>(scratch4, width, height, chroma_n, cur_y != nullptr, chroma, out4);
    VP8_LAUNCH_CHECK();
}
