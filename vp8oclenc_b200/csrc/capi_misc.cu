// Library identification entry points of libvp8b200.so.
#include <cstdio>
#include <cstring>

#include "common.cuh"

extern "C" const char *vp8b200_version(void) { return "vp8oclenc_b200 0.1 (sm_100a)"; }

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
