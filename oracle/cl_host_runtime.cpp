// oracle/cl_host_runtime.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A synchronous, host-memory implementation of the 28 OpenCL entry points of
// include/CL/cl.h.  Linked with ref_kernels_gpu.cpp / ref_kernels_cpu.cpp (the reference's
// own .cl kernels compiled for the CPU) it forms oracle/_ref/libOpenCL.so.1: the UNMODIFIED
// reference host (src/vp8enc.cpp + src/entropy_host.cpp) linked against it is the
// reference encoder running entirely on the CPU.
//
// Semantics fixed here: every enqueued command executes immediately, in
// program order -- a legal schedule of the host's in-order queues and what a
// single-threaded CPU OpenCL device would do.  Work-items of one NDRange run on OpenMP
// threads; the reference's kernels never communicate between work-items (their __local
// arrays are indexed by the work-item's own local id only), so that is exact.
#include <CL/cl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "clc_compat.hpp"

namespace clc {
thread_local work_item_state g_wi;
}

// ---- optional trace (VP8CL_TRACE=<file>): every create / write / read / image write is
// appended as {u32 kind, u32 mem index, u64 offset, u64 size, payload}.  Tests use it to
// replay the real host's per-frame inputs through other implementations and to compare
// what the host read back.  kind: 0 create buffer, 1 create image, 2 write, 3 read, 4 image write
static FILE *g_trace = nullptr;
static unsigned g_next_mem_index = 0;
static void trace_rec(unsigned kind, unsigned idx, size_t off, size_t size, const void *payload) {
    if (!g_trace) return;
    unsigned long long hdr[2] = {(unsigned long long)off, (unsigned long long)size};
    fwrite(&kind, 4, 1, g_trace);
    fwrite(&idx, 4, 1, g_trace);
    fwrite(hdr, 8, 2, g_trace);
    if (payload) fwrite(payload, 1, size, g_trace);
}
static void trace_init() {
    static bool done = false;
    if (done) return;
    done = true;
    const char *p = getenv("VP8CL_TRACE");
    if (p && *p) g_trace = fopen(p, "wb");
}

// ---- host-gap profiler (VP8CL_GAPS=<file>): where does the HOST PROGRAM spend its own time? ---------------
// The time between the return of one entry point and the call of the next is the host program's own work.  It is
// accumulated per (previous call, next call) pair, labels = entry point + object index / kernel name, and the
// largest pairs are written at exit.  The host is the same code whatever runtime it is linked against, so this
// profile (taken on the CPU runtime) says what the unmodified host costs per frame next to the CUDA shim.
#include <time.h>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
static bool g_gaps_on = false;
static std::string g_gap_prev = "start";
static unsigned long long g_gap_t = 0, g_gap_frames = 0;
static std::map<std::string, std::pair<unsigned long long, unsigned long long>> g_gaps;
static unsigned long long gap_now() {
    timespec ts;
    clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + ts.tv_nsec;
}
static void gaps_write() {
    const char *p = getenv("VP8CL_GAPS");
    FILE *f = p ? fopen(p, "w") : nullptr;
    if (!f) return;
    std::vector<std::pair<unsigned long long, std::string>> v;
    unsigned long long total = 0;
    for (auto &kv : g_gaps) { v.push_back({kv.second.first, kv.first}); total += kv.second.first; }
    std::sort(v.rbegin(), v.rend());
    fprintf(f, "host program between OpenCL calls: %.3f ms total over %llu inter frames (+ key frames)\n", total * 1e-6, g_gap_frames);
    for (size_t i = 0; i < v.size() && i < 40; ++i)
        fprintf(f, "%10.3f ms %8llu x  %s\n", v[i].first * 1e-6, g_gaps[v[i].second].second, v[i].second.c_str());
    fclose(f);
}
struct GapScope {  // at entry: close the gap; at exit: start the next one
    const char *fn;
    std::string label;
    GapScope(const char *f, const char *what, long idx) : fn(f) {
        if (!g_gaps_on) return;
        char b[96];
        if (what) snprintf(b, sizeof(b), "%s(%s)", f, what); else if (idx >= 0) snprintf(b, sizeof(b), "%s(#%ld)", f, idx); else snprintf(b, sizeof(b), "%s", f);
        label = b;
        const unsigned long long t = gap_now();
        auto &g = g_gaps[g_gap_prev + " -> " + label];
        g.first += t - g_gap_t;
        g.second += 1;
    }
    ~GapScope() {
        if (!g_gaps_on) return;
        g_gap_prev = label;
        g_gap_t = gap_now();
    }
};
static void gaps_init() {
    static bool done = false;
    if (done) return;
    done = true;
    if (getenv("VP8CL_GAPS")) { g_gaps_on = true; g_gap_t = gap_now(); atexit(gaps_write); }
}

// ---- per-stage profiler (VP8CL_STAGES=<file>): wall time of the reference's kernels by name --------------------
// Written at exit as "name calls wall_ms" lines plus the process' total wall time; bench.py groups them into the
// stages of BASELINE.md (motion search, predict + transform, loop filter, entropy; host = the rest).
static bool g_stages_on = false;
static unsigned long long g_stage_t0 = 0;
static std::map<std::string, std::pair<unsigned long long, unsigned long long>> g_stage_ns;
static unsigned long long wall_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + ts.tv_nsec;
}
static void stages_write() {
    const char *p = getenv("VP8CL_STAGES");
    FILE *f = p ? fopen(p, "w") : nullptr;
    if (!f) return;
    fprintf(f, "total_wall %.3f\n", (wall_now() - g_stage_t0) * 1e-6);
    for (auto &kv : g_stage_ns) fprintf(f, "%s %llu %.3f\n", kv.first.c_str(), kv.second.second, kv.second.first * 1e-6);
    fclose(f);
}
static void stages_init() {
    static bool done = false;
    if (done) return;
    done = true;
    if (getenv("VP8CL_STAGES")) { g_stages_on = true; g_stage_t0 = wall_now(); atexit(stages_write); }
}
struct StageScope {
    const char *name;
    unsigned long long t;
    explicit StageScope(const char *n) : name(n), t(g_stages_on ? wall_now() : 0) {}
    ~StageScope() {
        if (!g_stages_on) return;
        auto &e = g_stage_ns[name];
        e.first += wall_now() - t;
        e.second += 1;
    }
};

// ---- crash report: a SIGSEGV of the reference (host code or kernels on the CPU) prints where it happened ----------
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static const char *volatile g_running_kernel = nullptr;
static void crash_report(int sig, siginfo_t *info, void *) {
    char line[256];
    const char *k = g_running_kernel;
    const int n = snprintf(line, sizeof(line), "vp8ref runtime: signal %d at address %p, kernel %s\n", sig, info ? info->si_addr : nullptr, k ? k : "(host code)");
    if (n > 0) (void)!write(2, line, (size_t)n);
    void *frames[48];
    backtrace_symbols_fd(frames, backtrace(frames, 48), 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
static void crash_report_init() {
    static bool done = false;
    if (done) return;
    done = true;
    struct sigaction sa = {};
    sa.sa_sigaction = crash_report;
    sa.sa_flags = SA_SIGINFO;
    sigaction(SIGSEGV, &sa, nullptr);
    sigaction(SIGBUS, &sa, nullptr);
}

extern "C" const clc::kernel_desc vp8ref_gpu_kernels[];
extern "C" const clc::kernel_desc vp8ref_cpu_kernels[];

struct _cl_platform_id { int dummy; };
struct _cl_device_id { cl_device_type type; const char *name; };
struct _cl_context { cl_device_id dev; };
struct _cl_command_queue { cl_context ctx; };
struct _cl_mem {
    unsigned index;
    bool is_image;
    size_t size;
    unsigned char *data;
    clc::image2d img;
};
// The reference's search kernels read the previous frame WITHOUT clamping the candidate position (candidates that
// hang over the edge can never win, so what they read does not matter; src/GPU_kernels.cl:459-560): a few rows before
// and after the plane.  On a GPU that lands in neighbouring allocations; here a large plane is its own mmap and the
// read can hit an unmapped page -- the 2160p encode died with SIGSEGV about once in eight runs (luma_search_1step).
// Every object therefore gets zero-filled slack on both sides.
static const size_t kSlack = 1u << 20;
static unsigned char *alloc_with_slack(size_t size) {
    unsigned char *base = (unsigned char *)calloc(size + 2 * kSlack, 1);
    return base ? base + kSlack : nullptr;
}
static void free_with_slack(unsigned char *data) {
    if (data) free(data - kSlack);
}
struct _cl_program { bool is_gpu_program; };
struct _cl_kernel {
    const clc::kernel_desc *desc;
    unsigned char bytes[24][16];
    size_t sizes[24];
};

static _cl_platform_id g_platform;
static _cl_device_id g_cpu_dev = {CL_DEVICE_TYPE_CPU, "vp8oclenc reference kernels on host CPU (cpu device)"};
static _cl_device_id g_gpu_dev = {CL_DEVICE_TYPE_GPU, "vp8oclenc reference kernels on host CPU (gpu device)"};

static cl_int put_info(const void *src, size_t n, size_t cap, void *dst, size_t *ret) {
    if (ret) *ret = n;
    if (dst) {
        if (cap < n) return CL_INVALID_VALUE;
        memcpy(dst, src, n);
    }
    return CL_SUCCESS;
}

extern "C" {

cl_int clGetPlatformIDs(cl_uint num_entries, cl_platform_id *platforms, cl_uint *num_platforms) {
    gaps_init();
    stages_init();
    crash_report_init();
    if (num_platforms) *num_platforms = 1;
    if (platforms && num_entries >= 1) platforms[0] = &g_platform;
    return CL_SUCCESS;
}

cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t cap, void *dst, size_t *ret) {
    static const char name[] = "vp8oclenc reference-on-CPU platform (oracle/_ref)";
    return put_info(name, sizeof(name), cap, dst, ret);
}

cl_int clGetDeviceIDs(cl_platform_id, cl_device_type type, cl_uint num_entries, cl_device_id *devices,
                      cl_uint *num_devices) {
    cl_device_id found[2];
    cl_uint n = 0;
    if (type & CL_DEVICE_TYPE_CPU) found[n++] = &g_cpu_dev;
    if (type & CL_DEVICE_TYPE_GPU) found[n++] = &g_gpu_dev;
    if (num_devices) *num_devices = n;
    if (n == 0) return CL_DEVICE_NOT_FOUND;
    for (cl_uint i = 0; devices && i < n && i < num_entries; ++i) devices[i] = found[i];
    return CL_SUCCESS;
}

cl_int clGetDeviceInfo(cl_device_id dev, cl_device_info what, size_t cap, void *dst, size_t *ret) {
    switch (what) {
        case CL_DEVICE_NAME: return put_info(dev->name, strlen(dev->name) + 1, cap, dst, ret);
        case CL_DEVICE_VERSION: return put_info("OpenCL 1.1 subset", 18, cap, dst, ret);
        case CL_DRIVER_VERSION: return put_info("oracle/_ref", 12, cap, dst, ret);
        case CL_DEVICE_OPENCL_C_VERSION: return put_info("OpenCL C 1.0 (g++)", 19, cap, dst, ret);
        case CL_DEVICE_MAX_COMPUTE_UNITS: { cl_uint v = 1; return put_info(&v, sizeof(v), cap, dst, ret); }
        case CL_DEVICE_MAX_WORK_GROUP_SIZE: { size_t v = 256; return put_info(&v, cap < sizeof(v) ? cap : sizeof(v), cap, dst, ret); }
        case CL_DEVICE_TYPE: return put_info(&dev->type, sizeof(dev->type), cap, dst, ret);
        default: return CL_INVALID_VALUE;
    }
}

cl_context clCreateContext(const cl_context_properties *, cl_uint n, const cl_device_id *devs,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *err) {
    if (err) *err = (n >= 1 && devs) ? CL_SUCCESS : CL_INVALID_VALUE;
    if (n < 1 || !devs) return nullptr;
    return new _cl_context{devs[0]};
}
cl_int clReleaseContext(cl_context c) { delete c; return CL_SUCCESS; }

cl_command_queue clCreateCommandQueue(cl_context ctx, cl_device_id, cl_command_queue_properties, cl_int *err) {
    if (err) *err = CL_SUCCESS;
    return new _cl_command_queue{ctx};
}
cl_int clReleaseCommandQueue(cl_command_queue q) { delete q; return CL_SUCCESS; }

cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t size, void *host_ptr, cl_int *err) {
    trace_init();
    _cl_mem *m = new _cl_mem();
    m->index = g_next_mem_index++;
    m->is_image = false;
    m->size = size;
    trace_rec(0, m->index, 0, size, nullptr);
    m->data = alloc_with_slack(size);
    if (host_ptr && m->data) memcpy(m->data, host_ptr, size);
    if (err) *err = m->data ? CL_SUCCESS : CL_MEM_OBJECT_ALLOCATION_FAILURE;
    return m;
}

cl_mem clCreateImage2D(cl_context, cl_mem_flags, const cl_image_format *fmt, size_t w, size_t h, size_t,
                       void *, cl_int *err) {
    if (!fmt || fmt->image_channel_order != CL_R || fmt->image_channel_data_type != CL_UNSIGNED_INT8) {
        if (err) *err = CL_INVALID_IMAGE_FORMAT_DESCRIPTOR;
        return nullptr;
    }
    trace_init();
    _cl_mem *m = new _cl_mem();
    m->index = g_next_mem_index++;
    m->is_image = true;
    m->size = w * h;
    trace_rec(1, m->index, w, h, nullptr);
    m->data = alloc_with_slack(m->size);
    m->img.data = m->data;
    m->img.width = (int)w;
    m->img.height = (int)h;
    if (err) *err = CL_SUCCESS;
    return m;
}

cl_int clReleaseMemObject(cl_mem m) {
    if (!m) return CL_INVALID_MEM_OBJECT;
    free_with_slack(m->data);
    delete m;
    return CL_SUCCESS;
}

cl_program clCreateProgramWithSource(cl_context, cl_uint count, const char **strings, const size_t *, cl_int *err) {
    bool gpu = false;
    for (cl_uint i = 0; i < count; ++i)
        if (strings[i] && strstr(strings[i], "luma_search_1step")) gpu = true;
    if (err) *err = CL_SUCCESS;
    return new _cl_program{gpu};
}
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *, void (*)(cl_program, void *), void *) {
    return CL_SUCCESS;
}
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t cap, void *dst, size_t *ret) {
    return put_info("", 1, cap, dst, ret);
}
cl_int clReleaseProgram(cl_program p) { delete p; return CL_SUCCESS; }

cl_kernel clCreateKernel(cl_program prog, const char *name, cl_int *err) {
    const clc::kernel_desc *tab = prog->is_gpu_program ? vp8ref_gpu_kernels : vp8ref_cpu_kernels;
    for (; tab->name; ++tab)
        if (!strcmp(tab->name, name)) {
            _cl_kernel *k = new _cl_kernel();
            memset(k, 0, sizeof(*k));
            k->desc = tab;
            if (err) *err = CL_SUCCESS;
            return k;
        }
    if (err) *err = CL_INVALID_KERNEL_NAME;
    return nullptr;
}
cl_int clReleaseKernel(cl_kernel k) { delete k; return CL_SUCCESS; }

cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value) {
    GapScope gs("setarg", nullptr, -1);
    if (!k) return CL_INVALID_KERNEL;
    if ((int)idx >= k->desc->nargs) return CL_INVALID_ARG_INDEX;
    if (size > 16) return CL_INVALID_ARG_SIZE;
    memset(k->bytes[idx], 0, 16);
    if (value) memcpy(k->bytes[idx], value, size);
    k->sizes[idx] = size;
    return CL_SUCCESS;
}

cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel k, cl_uint dim, const size_t *, const size_t *gsz,
                              const size_t *lsz, cl_uint, const cl_event *, cl_event *) {
    GapScope gs("kernel", k ? k->desc->name : "?", -1); if (g_gaps_on && k && !strcmp(k->desc->name, "reset_vectors")) ++g_gap_frames;
    if (!k) return CL_INVALID_KERNEL;
    if (dim != 1) return CL_INVALID_WORK_DIMENSION;
    StageScope stage_scope(k->desc->name);
    g_running_kernel = k->desc->name;
    struct KernelNameScope { ~KernelNameScope() { g_running_kernel = nullptr; } } kernel_name_scope;
    const int n = k->desc->nargs;
    void *resolved[24];
    void *argv[24];
    for (int i = 0; i < n; ++i) {
        if (k->desc->is_pointer[i]) {
            cl_mem m;
            memcpy(&m, k->bytes[i], sizeof(m));
            resolved[i] = !m ? nullptr : (m->is_image ? (void *)&m->img : (void *)m->data);
            argv[i] = &resolved[i];
        } else {
            argv[i] = k->bytes[i];
        }
    }
    const long long global = (long long)gsz[0];
    const long long local = lsz ? (long long)lsz[0] : 1;
    const clc::kernel_desc *d = k->desc;
#pragma omp parallel for schedule(dynamic, 1) if (global >= 64)
    for (long long g0 = 0; g0 < global; g0 += 64) {
        const long long g1 = g0 + 64 < global ? g0 + 64 : global;
        for (long long g = g0; g < g1; ++g) {
            clc::g_wi.global_id = (size_t)g;
            clc::g_wi.global_size = (size_t)global;
            clc::g_wi.local_size = (size_t)local;
            clc::g_wi.local_id = (size_t)(g % local);
            clc::g_wi.group_id = (size_t)(g / local);
            d->invoke(argv);
        }
    }
    return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t size, void *ptr, cl_uint,
                           const cl_event *, cl_event *) {
    GapScope gs("read", nullptr, m ? (long)m->index : -1);
    if (!m || off + size > m->size) return CL_INVALID_VALUE;
    memcpy(ptr, m->data + off, size);
    trace_rec(3, m->index, off, size, ptr);
    return CL_SUCCESS;
}
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem m, cl_bool, size_t off, size_t size, const void *ptr, cl_uint,
                            const cl_event *, cl_event *) {
    GapScope gs("write", nullptr, m ? (long)m->index : -1);
    if (!m || off + size > m->size) return CL_INVALID_VALUE;
    memcpy(m->data + off, ptr, size);
    trace_rec(2, m->index, off, size, ptr);
    return CL_SUCCESS;
}
cl_int clEnqueueCopyBuffer(cl_command_queue, cl_mem s, cl_mem d, size_t so, size_t dof, size_t size, cl_uint,
                           const cl_event *, cl_event *) {
    GapScope gs("copy", nullptr, -1);
    if (!s || !d || so + size > s->size || dof + size > d->size) return CL_INVALID_VALUE;
    memmove(d->data + dof, s->data + so, size);
    return CL_SUCCESS;
}
cl_int clEnqueueWriteImage(cl_command_queue, cl_mem img, cl_bool, const size_t *origin, const size_t *region,
                           size_t row_pitch, size_t, const void *ptr, cl_uint, const cl_event *, cl_event *) {
    GapScope gs("writeimage", nullptr, img ? (long)img->index : -1);
    if (!img || !img->is_image) return CL_INVALID_MEM_OBJECT;
    const size_t pitch = row_pitch ? row_pitch : region[0];
    for (size_t y = 0; y < region[1]; ++y)
        memcpy(img->data + (origin[1] + y) * img->img.width + origin[0], (const unsigned char *)ptr + y * pitch,
               region[0]);
    if (!row_pitch || row_pitch == region[0]) trace_rec(4, img->index, 0, region[0] * region[1], ptr);
    return CL_SUCCESS;
}
cl_int clEnqueueCopyImage(cl_command_queue, cl_mem s, cl_mem d, const size_t *so, const size_t *dor,
                          const size_t *region, cl_uint, const cl_event *, cl_event *) {
    GapScope gs("copyimage", nullptr, -1);
    if (!s || !d || !s->is_image || !d->is_image) return CL_INVALID_MEM_OBJECT;
    for (size_t y = 0; y < region[1]; ++y)
        memmove(d->data + (dor[1] + y) * d->img.width + dor[0], s->data + (so[1] + y) * s->img.width + so[0],
                region[0]);
    return CL_SUCCESS;
}
void *clEnqueueMapBuffer(cl_command_queue, cl_mem m, cl_bool, cl_map_flags, size_t off, size_t size, cl_uint,
                         const cl_event *, cl_event *, cl_int *err) {
    GapScope gs("map", nullptr, m ? (long)m->index : -1);
    if (!m || off + size > m->size) {
        if (err) *err = CL_INVALID_VALUE;
        return nullptr;
    }
    if (err) *err = CL_SUCCESS;
    return m->data + off;
}
cl_int clEnqueueUnmapMemObject(cl_command_queue, cl_mem m, void *, cl_uint, const cl_event *, cl_event *) {
    GapScope gs("unmap", nullptr, m ? (long)m->index : -1);
    if (m) trace_rec(5, m->index, 0, m->size, m->data); /* kind 5: contents handed back by the host */
    return CL_SUCCESS;
}
cl_int clFlush(cl_command_queue) {
    GapScope gs("flush", nullptr, -1);
    return CL_SUCCESS;
}
cl_int clFinish(cl_command_queue) {
    GapScope gs("finish", nullptr, -1);
    if (g_trace) fflush(g_trace);
    return CL_SUCCESS;
}

// ---- test hook: run one reference kernel by name on raw host pointers ----------------
// argv[i] points at the argument value: for pointer parameters a void* holding the host
// address (for image parameters: the address of a {data,width,height} clc::image2d).
int vp8ref_run_kernel(const char *name, long long global, long long local, void *const *argv) {
    const clc::kernel_desc *tabs[2] = {vp8ref_gpu_kernels, vp8ref_cpu_kernels};
    for (int t = 0; t < 2; ++t)
        for (const clc::kernel_desc *d = tabs[t]; d->name; ++d)
            if (!strcmp(d->name, name)) {
                if (local < 1) local = 1;
#pragma omp parallel for schedule(dynamic, 1) if (global >= 64)
                for (long long g0 = 0; g0 < global; g0 += 64) {
                    const long long g1 = g0 + 64 < global ? g0 + 64 : global;
                    for (long long g = g0; g < g1; ++g) {
                        clc::g_wi.global_id = (size_t)g;
                        clc::g_wi.global_size = (size_t)global;
                        clc::g_wi.local_size = (size_t)local;
                        clc::g_wi.local_id = (size_t)(g % local);
                        clc::g_wi.group_id = (size_t)(g / local);
                        d->invoke(argv);
                    }
                }
                return 0;
            }
    return -1;
}
int vp8ref_kernel_nargs(const char *name) {
    const clc::kernel_desc *tabs[2] = {vp8ref_gpu_kernels, vp8ref_cpu_kernels};
    for (int t = 0; t < 2; ++t)
        for (const clc::kernel_desc *d = tabs[t]; d->name; ++d)
            if (!strcmp(d->name, name)) return d->nargs;
    return -1;
}

}  // extern "C"
