// Library identification entry points of libvp8b200.so.
#include <cstdio>
#include <cstring>

#include "common.cuh"

extern "C" const char *vp8b200_version(void) { return "vp8oclenc_b200 0.1 (sm_100a)"; }

// ---- integer-issue micro-benchmark: the roofline denominator of the motion-search kernels ----
// The search kernels execute a mix of IMAD (fma pipe) and IADD3/LOP3/SHF (alu pipe) on 32-bit
// integers.  This kernel issues the same two classes in a 1:1 ratio from 8 independent
// dependency chains per thread, enough resident warps to saturate both pipes; ops counted:
// one per IMAD, one per add/logic instruction.
__global__ void __launch_bounds__(256) k_int_peak(int *out, int iters, int seed) {
    int a0 = threadIdx.x + seed, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 + 11, a5 = a0 + 13, a6 = a0 ^ 17, a7 = a0 ^ 19;
    const int m = seed | 1, c = seed + 12345;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = a0 * m + c;  a1 = (a1 ^ a0) + c;
            a2 = a2 * m + c;  a3 = (a3 ^ a2) + c;
            a4 = a4 * m + c;  a5 = (a5 ^ a4) + c;
            a6 = a6 * m + c;  a7 = (a7 ^ a6) + c;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// runs the micro-benchmark on `stream`; returns the measured 32-bit integer ops per second
extern "C" double vp8b200_measure_int_ops_per_second(void *stream, int repeats) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    int *out = nullptr;
    if (cudaMalloc((void **)&out, (size_t)blocks * threads * sizeof(int)) != cudaSuccess) return -1.0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_int_peak<<<blocks, threads, 0, st>>>(out, 64, 1);  // warm-up
    double best = 0.0;
    for (int r = 0; r < repeats; ++r) {
        cudaEventRecord(e0, st);
        k_int_peak<<<blocks, threads, 0, st>>>(out, iters, r + 2);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // per inner step: 4 IMAD + 4 LOP3 + 4 IADD = 12 integer instructions per thread... counted as
        // 4 multiply-adds + 8 add/logic = 12 ops
        const double ops = (double)blocks * threads * (double)iters * 8.0 * 12.0;
        if (ms > 0.f) best = ops / (ms * 1e-3) > best ? ops / (ms * 1e-3) : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

extern "C" int vp8b200_device_info(char *name, int name_cap, int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return -(int)e;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return -(int)e;
    if (name && name_cap > 0) {
        strncpy(name, p.name, (size_t)name_cap - 1);
        name[name_cap - 1] = 0;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return 0;
}
