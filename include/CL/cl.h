/*
 * CL/cl.h -- the OpenCL 1.1 C-API subset that vp8oclenc's host code calls.
 *
 * This is the drop-in boundary of vp8oclenc_b200.  The
 * reference host (src/vp8enc.cpp, src/init.h, src/inter_part.h,
 * src/loop_filter.h, src/intra_part.h, src/encIO.h) includes <CL/cl.h>
 * (src/vp8enc.h:1) and links with -lOpenCL ("makefile example":2).  It uses
 * exactly the 28 entry points declared at the bottom of this file, the scalar
 * typedefs, cl_image_format and the CL_* constants below; nothing else of
 * OpenCL is needed, so nothing else is declared.
 *
 * Two libraries implement this header in the repo:
 *   - vp8oclenc_b200/lib/libOpenCL.so.1 : the product.  Every kernel of the
 *     "GPU program" and the loop filter of the "CPU program" run as
 *     hand-written CUDA on a B200 (vp8oclenc_b200/csrc).
 *   - oracle/_ref/libOpenCL.so.1        : test infrastructure.  The
 *     reference's own .cl kernels compiled for the host CPU (oracle/Makefile).
 * The numeric values of the constants follow the Khronos OpenCL 1.1
 * specification, so a host binary built against a vendor's cl.h also works.
 */
#ifndef VP8B200_CL_CL_H
#define VP8B200_CL_CL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* scalar types */
typedef int8_t   cl_char;
typedef uint8_t  cl_uchar;
typedef int16_t  cl_short;
typedef uint16_t cl_ushort;
typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef int64_t  cl_long;
typedef uint64_t cl_ulong;
typedef float    cl_float;
typedef double   cl_double;

/* opaque handles */
typedef struct _cl_platform_id   *cl_platform_id;
typedef struct _cl_device_id     *cl_device_id;
typedef struct _cl_context       *cl_context;
typedef struct _cl_command_queue *cl_command_queue;
typedef struct _cl_mem           *cl_mem;
typedef struct _cl_program       *cl_program;
typedef struct _cl_kernel        *cl_kernel;
typedef struct _cl_event         *cl_event;
typedef struct _cl_sampler       *cl_sampler;

typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_uint  cl_platform_info;
typedef cl_uint  cl_device_info;
typedef cl_bitfield cl_command_queue_properties;
typedef intptr_t cl_context_properties;
typedef cl_bitfield cl_mem_flags;
typedef cl_uint  cl_channel_order;
typedef cl_uint  cl_channel_type;
typedef cl_bitfield cl_map_flags;
typedef cl_uint  cl_program_build_info;

typedef struct _cl_image_format {
    cl_channel_order image_channel_order;
    cl_channel_type  image_channel_data_type;
} cl_image_format;

/* error codes */
#define CL_SUCCESS                         0
#define CL_DEVICE_NOT_FOUND               -1
#define CL_MEM_OBJECT_ALLOCATION_FAILURE  -4
#define CL_OUT_OF_RESOURCES               -5
#define CL_OUT_OF_HOST_MEMORY             -6
#define CL_BUILD_PROGRAM_FAILURE          -11
#define CL_MAP_FAILURE                    -12
#define CL_INVALID_VALUE                  -30
#define CL_INVALID_DEVICE_TYPE            -31
#define CL_INVALID_PLATFORM               -32
#define CL_INVALID_DEVICE                 -33
#define CL_INVALID_CONTEXT                -34
#define CL_INVALID_COMMAND_QUEUE          -36
#define CL_INVALID_MEM_OBJECT             -38
#define CL_INVALID_IMAGE_FORMAT_DESCRIPTOR -39
#define CL_INVALID_PROGRAM                -44
#define CL_INVALID_PROGRAM_EXECUTABLE     -45
#define CL_INVALID_KERNEL_NAME            -46
#define CL_INVALID_KERNEL                 -48
#define CL_INVALID_ARG_INDEX              -49
#define CL_INVALID_ARG_VALUE              -50
#define CL_INVALID_ARG_SIZE               -51
#define CL_INVALID_KERNEL_ARGS            -52
#define CL_INVALID_WORK_DIMENSION         -53
#define CL_INVALID_WORK_GROUP_SIZE        -54
#define CL_INVALID_GLOBAL_WORK_SIZE       -63

#define CL_FALSE 0
#define CL_TRUE  1

/* cl_platform_info */
#define CL_PLATFORM_PROFILE  0x0900
#define CL_PLATFORM_VERSION  0x0901
#define CL_PLATFORM_NAME     0x0902
#define CL_PLATFORM_VENDOR   0x0903

/* cl_device_type */
#define CL_DEVICE_TYPE_DEFAULT     (1 << 0)
#define CL_DEVICE_TYPE_CPU         (1 << 1)
#define CL_DEVICE_TYPE_GPU         (1 << 2)
#define CL_DEVICE_TYPE_ACCELERATOR (1 << 3)
#define CL_DEVICE_TYPE_ALL         0xFFFFFFFF

/* cl_device_info */
#define CL_DEVICE_TYPE                 0x1000
#define CL_DEVICE_MAX_COMPUTE_UNITS    0x1002
#define CL_DEVICE_MAX_WORK_GROUP_SIZE  0x1004
#define CL_DEVICE_NAME                 0x102B
#define CL_DEVICE_VENDOR               0x102C
#define CL_DRIVER_VERSION              0x102D
#define CL_DEVICE_VERSION              0x102F
#define CL_DEVICE_OPENCL_C_VERSION     0x103D

/* cl_mem_flags */
#define CL_MEM_READ_WRITE      (1 << 0)
#define CL_MEM_WRITE_ONLY      (1 << 1)
#define CL_MEM_READ_ONLY       (1 << 2)
#define CL_MEM_USE_HOST_PTR    (1 << 3)
#define CL_MEM_ALLOC_HOST_PTR  (1 << 4)
#define CL_MEM_COPY_HOST_PTR   (1 << 5)

/* image formats */
#define CL_R               0x10B0
#define CL_UNSIGNED_INT8   0x10DA

/* cl_map_flags */
#define CL_MAP_READ                     (1 << 0)
#define CL_MAP_WRITE                    (1 << 1)
#define CL_MAP_WRITE_INVALIDATE_REGION  (1 << 2)

/* cl_program_build_info */
#define CL_PROGRAM_BUILD_STATUS   0x1181
#define CL_PROGRAM_BUILD_OPTIONS  0x1182
#define CL_PROGRAM_BUILD_LOG      0x1183

#define CL_API_ENTRY
#define CL_API_CALL
#define CL_CALLBACK

/* ---- the 28 entry points (callers: src/init.h:23-100,102-1278;
 *      src/vp8enc.cpp:48-94,224-227,354-470,501-708; src/inter_part.h:1-384;
 *      src/loop_filter.h:1-190; src/intra_part.h:1114-1125; src/encIO.h:4-27) ---- */

cl_int clGetPlatformIDs(cl_uint num_entries, cl_platform_id *platforms, cl_uint *num_platforms);
cl_int clGetPlatformInfo(cl_platform_id platform, cl_platform_info param_name,
                         size_t param_value_size, void *param_value, size_t *param_value_size_ret);
cl_int clGetDeviceIDs(cl_platform_id platform, cl_device_type device_type, cl_uint num_entries,
                      cl_device_id *devices, cl_uint *num_devices);
cl_int clGetDeviceInfo(cl_device_id device, cl_device_info param_name,
                       size_t param_value_size, void *param_value, size_t *param_value_size_ret);

cl_context clCreateContext(const cl_context_properties *properties, cl_uint num_devices,
                           const cl_device_id *devices,
                           void (CL_CALLBACK *pfn_notify)(const char *, const void *, size_t, void *),
                           void *user_data, cl_int *errcode_ret);
cl_int clReleaseContext(cl_context context);

cl_command_queue clCreateCommandQueue(cl_context context, cl_device_id device,
                                      cl_command_queue_properties properties, cl_int *errcode_ret);
cl_int clReleaseCommandQueue(cl_command_queue command_queue);

cl_mem clCreateBuffer(cl_context context, cl_mem_flags flags, size_t size, void *host_ptr,
                      cl_int *errcode_ret);
cl_mem clCreateImage2D(cl_context context, cl_mem_flags flags, const cl_image_format *image_format,
                       size_t image_width, size_t image_height, size_t image_row_pitch,
                       void *host_ptr, cl_int *errcode_ret);
cl_int clReleaseMemObject(cl_mem memobj);

cl_program clCreateProgramWithSource(cl_context context, cl_uint count, const char **strings,
                                     const size_t *lengths, cl_int *errcode_ret);
cl_int clBuildProgram(cl_program program, cl_uint num_devices, const cl_device_id *device_list,
                      const char *options,
                      void (CL_CALLBACK *pfn_notify)(cl_program, void *), void *user_data);
cl_int clGetProgramBuildInfo(cl_program program, cl_device_id device,
                             cl_program_build_info param_name, size_t param_value_size,
                             void *param_value, size_t *param_value_size_ret);
cl_int clReleaseProgram(cl_program program);

cl_kernel clCreateKernel(cl_program program, const char *kernel_name, cl_int *errcode_ret);
cl_int clReleaseKernel(cl_kernel kernel);
cl_int clSetKernelArg(cl_kernel kernel, cl_uint arg_index, size_t arg_size, const void *arg_value);

cl_int clEnqueueNDRangeKernel(cl_command_queue command_queue, cl_kernel kernel, cl_uint work_dim,
                              const size_t *global_work_offset, const size_t *global_work_size,
                              const size_t *local_work_size, cl_uint num_events_in_wait_list,
                              const cl_event *event_wait_list, cl_event *event);
cl_int clEnqueueReadBuffer(cl_command_queue command_queue, cl_mem buffer, cl_bool blocking_read,
                           size_t offset, size_t size, void *ptr, cl_uint num_events_in_wait_list,
                           const cl_event *event_wait_list, cl_event *event);
cl_int clEnqueueWriteBuffer(cl_command_queue command_queue, cl_mem buffer, cl_bool blocking_write,
                            size_t offset, size_t size, const void *ptr,
                            cl_uint num_events_in_wait_list, const cl_event *event_wait_list,
                            cl_event *event);
cl_int clEnqueueCopyBuffer(cl_command_queue command_queue, cl_mem src_buffer, cl_mem dst_buffer,
                           size_t src_offset, size_t dst_offset, size_t size,
                           cl_uint num_events_in_wait_list, const cl_event *event_wait_list,
                           cl_event *event);
cl_int clEnqueueWriteImage(cl_command_queue command_queue, cl_mem image, cl_bool blocking_write,
                           const size_t *origin, const size_t *region, size_t input_row_pitch,
                           size_t input_slice_pitch, const void *ptr,
                           cl_uint num_events_in_wait_list, const cl_event *event_wait_list,
                           cl_event *event);
cl_int clEnqueueCopyImage(cl_command_queue command_queue, cl_mem src_image, cl_mem dst_image,
                          const size_t *src_origin, const size_t *dst_origin, const size_t *region,
                          cl_uint num_events_in_wait_list, const cl_event *event_wait_list,
                          cl_event *event);
void *clEnqueueMapBuffer(cl_command_queue command_queue, cl_mem buffer, cl_bool blocking_map,
                         cl_map_flags map_flags, size_t offset, size_t size,
                         cl_uint num_events_in_wait_list, const cl_event *event_wait_list,
                         cl_event *event, cl_int *errcode_ret);
cl_int clEnqueueUnmapMemObject(cl_command_queue command_queue, cl_mem memobj, void *mapped_ptr,
                               cl_uint num_events_in_wait_list, const cl_event *event_wait_list,
                               cl_event *event);
cl_int clFlush(cl_command_queue command_queue);
cl_int clFinish(cl_command_queue command_queue);

#ifdef __cplusplus
}
#endif

#endif /* VP8B200_CL_CL_H */
