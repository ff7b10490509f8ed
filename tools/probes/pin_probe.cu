// probe: is an H2D copy out of a malloc'ed buffer still right after a page-rounded part of it was cudaHostRegister'ed?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
int main() {
    const size_t luma = 352 * 288, full = luma * 3 / 2;
    unsigned char *pack = (unsigned char *)malloc(full), *back = (unsigned char *)malloc(luma), *dev;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaMalloc(&dev, luma);
    FILE *f = fopen("/dev/shm/pin_probe.bin", "wb");
    for (int fr = 0; fr < 6; ++fr) { for (size_t i = 0; i < full; ++i) pack[i] = (unsigned char)(i * 7 + fr * 31); fwrite(pack, 1, full, f); }
    fclose(f);
    f = fopen("/dev/shm/pin_probe.bin", "rb");
    for (int fr = 0; fr < 6; ++fr) {
        size_t got = fread(pack, 1, full, f);
        if (fr == 2) {
            const size_t page = sysconf(_SC_PAGESIZE), lo = (size_t)pack / page * page, hi = ((size_t)pack + luma + page - 1) / page * page;
            cudaError_t e = cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterDefault);
            printf("register %p..%p (buffer %p): %s\n", (void *)lo, (void *)hi, pack, cudaGetErrorString(e));
        }
        cudaMemcpyAsync(dev, pack, luma, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        cudaMemcpy(back, dev, luma, cudaMemcpyDeviceToHost);
        size_t bad = 0, first = 0;
        for (size_t i = 0; i < luma; ++i) if (back[i] != (unsigned char)(i * 7 + fr * 31)) { if (!bad) first = i; ++bad; }
        printf("frame %d: read %zu, %zu bad bytes (first %zu), host itself %s\n", fr, got, bad, first,
               pack[5] == (unsigned char)(5 * 7 + fr * 31) ? "ok" : "STALE");
    }
    return 0;
}
