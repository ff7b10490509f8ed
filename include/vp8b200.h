/*
 * vp8b200.h -- kernel-level C ABI of the B200 engine behind vp8oclenc's OpenCL boundary.
 *
 * One entry point per reference __kernel on the inter-frame hot path.
 * Each call has whole-NDRange semantics: it does what the reference host obtains by
 * enqueueing that kernel over its full global size (src/inter_part.h, src/loop_filter.h)
 * with the arguments bound in src/init.h:597-1271.  All pointers are DEVICE pointers;
 * `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are
 * asynchronous; the return value is 0 or a negative cudaError_t of the launch.
 *
 * The library (vp8oclenc_b200/lib/libvp8b200.so) contains only CUDA code for sm_100a: there
 * is no CPU fallback, every entry point fails with a CUDA error when no device is present.
 *
 * Layouts (identical to the reference's, src/vp8enc.h:80-120):
 *   planes        tightly packed uint8, luma stride = width, chroma = width/2
 *   vector nets   short2 per 8x8 block, stride net_width = 2*mb_width at EVERY pyramid level
 *   MB_vectors    4 x short2 per macroblock (TL,TR,BL,BR), quarter-pel units
 *   MB coeffs     25 blocks x 16 int16 per macroblock in zig-zag position order;
 *                 blocks 0-15 Y raster, 16-19 U, 20-23 V, 24 Y2
 *   segment_data  4 x 11 int32
 */
#ifndef VP8B200_H
#define VP8B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t y_ac_i, y_dc_idelta, y2_dc_idelta, y2_ac_idelta, uv_dc_idelta, uv_ac_idelta;
    int32_t loop_filter_level, mbedge_limit, sub_bedge_limit, interior_limit, hev_threshold;
} vp8b200_segment_data; /* src/vp8enc.h:80-92, src/GPU_kernels.cl:24-36 */

/* library / device identification; returns 0 when a CUDA device is usable */
int vp8b200_device_info(char *name, int name_cap, int *sm_count, int *cc_major, int *cc_minor);
const char *vp8b200_version(void);

/* Measures the 32-bit integer instruction rate of the device (ops/s) with a 1:2 IMAD : add/logic
 * micro-benchmark on `stream`; the roofline denominator of the motion-search kernels, which are
 * integer-ALU bound.  Returns < 0 on a CUDA error. */
double vp8b200_measure_int_ops_per_second(void *stream, int repeats);

/* replaces reset_vectors, src/GPU_kernels.cl:404-427 (enqueued at src/inter_part.h:5-6); n = 4*mb_count */
int vp8b200_reset_vectors(void *stream, int16_t *last_net1, int16_t *last_net2, int16_t *golden_net1,
                          int16_t *golden_net2, int16_t *altref_net1, int16_t *altref_net2, int32_t *last_Bdiff,
                          int32_t *golden_Bdiff, int32_t *altref_Bdiff, int n);

/* replaces downsample_x2, src/GPU_kernels.cl:429-451 (src/inter_part.h:11-33) */
int vp8b200_downsample_x2(void *stream, const uint8_t *src, uint8_t *dst, int src_width, int src_height);

/* replaces luma_search_1step, src/GPU_kernels.cl:459-560 (src/inter_part.h:109-219) */
int vp8b200_luma_search_1step(void *stream, const uint8_t *current_frame, const uint8_t *prev_frame,
                              const int16_t *src_net, int16_t *dst_net, int net_width, int width, int height,
                              int pixel_rate);

/* replaces luma_search_2step + construct_opt1/2, src/GPU_kernels.cl:776-1203 (src/inter_part.h:221-236).
 * ref_frame is the width x height luma plane the reference binds as an image (clamp-to-edge reads). */
int vp8b200_luma_search_2step(void *stream, const uint8_t *current_frame, const uint8_t *ref_frame,
                              const int16_t *net, int16_t *ref_net, int32_t *ref_Bdiff, int width, int height);

/* The same two kernels for up to three references (LAST, GOLDEN, ALTREF) in ONE launch: the host
 * enqueues the per-reference instances back to back on three queues (src/inter_part.h:122-135 and
 * the following levels); they are independent, so they run as grid.y = nrefs. */
int vp8b200_luma_search_1step_multi(void *stream, const uint8_t *current_frame, int nrefs,
                                    const uint8_t *const *prev_frame, const int16_t *const *src_net,
                                    int16_t *const *dst_net, int net_width, int width, int height, int pixel_rate);
int vp8b200_luma_search_2step_multi(void *stream, const uint8_t *current_frame, int nrefs,
                                    const uint8_t *const *ref_frame, const int16_t *const *net, int16_t *const *ref_net,
                                    int32_t *const *ref_Bdiff, int width, int height);

/* Experiment (not used by the frame pipeline): the same search with its reference windows staged by TMA 2-D box
 * copies out of a replicate-padded copy of the plane instead of word loads + funnel shifts.  `padded` is scratch of
 * (width + 32) * (height + 32) bytes; pad_only != 0 only builds the padded plane (to time the two parts apart).
 * Bit-identical results; the A/B numbers are in DESIGN.md section 4. */
int vp8b200_experiment_search_2step_tma(void *stream, const uint8_t *current_frame, const uint8_t *ref_frame, uint8_t *padded,
                                        const int16_t *net, int16_t *ref_net, int32_t *ref_Bdiff, int width, int height,
                                        int pad_only);

/* replaces select_reference, src/GPU_kernels.cl:1205-1283 (src/inter_part.h:250-255) */
int vp8b200_select_reference(void *stream, const int16_t *last_net, const int16_t *golden_net,
                             const int16_t *altref_net, const int32_t *last_Bdiff, const int32_t *golden_Bdiff,
                             const int32_t *altref_Bdiff, int32_t *MB_reference_frame, int16_t *MB_vectors, int width,
                             int height, int use_golden, int use_altref);

/* replaces pack_8x8_into_16x16, src/GPU_kernels.cl:1346-1366 (src/inter_part.h:257-258) */
int vp8b200_pack_8x8_into_16x16(void *stream, const int16_t *MB_vectors, int32_t *MB_parts, float *MB_SSIM,
                                int mb_count);

/* replaces prepare_predictors_and_residual + construct, src/GPU_kernels.cl:574-774,1285-1344
 * (src/inter_part.h:268-321).  width/height are the plane's; plane 0=Y 1=U 2=V; ref 0=LAST 1=GOLDEN 2=ALTREF */
int vp8b200_prepare_predictors_and_residual(void *stream, const uint8_t *current_frame, const uint8_t *ref_frame,
                                            uint8_t *predictor, int16_t *residual, const int32_t *MB_reference_frame,
                                            const int16_t *MB_vectors, int width, int height, int plane, int ref);

/* replaces dct4x4, src/GPU_kernels.cl:1368-1496 (src/inter_part.h:334-343) */
int vp8b200_dct4x4(void *stream, const int16_t *residual, int16_t *MB, int32_t *MB_segment_id,
                   const int32_t *MB_parts, const float *MB_SSIM, int width, int height,
                   const vp8b200_segment_data *SD, int segment_id, float SSIM_target, int plane);

/* replaces wht4x4_iwht4x4, src/GPU_kernels.cl:257-401,1498-1543 (src/inter_part.h:345-346) */
int vp8b200_wht4x4_iwht4x4(void *stream, int16_t *MB, int32_t *MB_segment_id, const int32_t *MB_parts,
                           const vp8b200_segment_data *SD, int segment_id, int mb_count);

/* replaces idct4x4, src/GPU_kernels.cl:192-255,1545-1608 (src/inter_part.h:349-357) */
int vp8b200_idct4x4(void *stream, uint8_t *recon_frame, const uint8_t *predictor, const int16_t *MB,
                    const int32_t *MB_segment_id, const int32_t *MB_parts, int width, int height,
                    const vp8b200_segment_data *SD, int segment_id, int plane);

/* replaces count_SSIM_luma (mb_size 16) and count_SSIM_chroma (mb_size 8),
 * src/GPU_kernels.cl:1610-2095 (src/inter_part.h:361-370) */
int vp8b200_count_SSIM(void *stream, const uint8_t *frame1, const uint8_t *frame2, const int32_t *MB_segment_id,
                       float *metric, int width, int height, int segment_id, int mb_size);

/* replaces gather_SSIM, src/GPU_kernels.cl:2097-2105 (src/inter_part.h:377) */
int vp8b200_gather_SSIM(void *stream, const float *metric1, const float *metric2, const float *metric3,
                        float *MB_SSIM, int mb_count);

/* ONE launch for the tail of inter_transform() (src/inter_part.h:268-378): replaces
 * prepare_predictors_and_residual x9, and per segment dct4x4 x3, wht4x4_iwht4x4, idct4x4 x3,
 * count_SSIM_luma, count_SSIM_chroma x2, gather_SSIM (53 launches).  img[3*ref + plane] are the
 * reference planes (NULL where a reference is unused).  Predictors/residuals stay on chip.
 * Requires SSIM_target >= -2 (true for every value the reference's CLI can produce). */
int vp8b200_mb_predict_transform_fused(void *stream, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                                       const uint8_t *const img[9], const int32_t *MB_reference_frame,
                                       const int16_t *MB_vectors, const int32_t *MB_parts, int16_t *MB,
                                       int32_t *MB_segment_id, float *MB_SSIM, uint8_t *recon_y, uint8_t *recon_u,
                                       uint8_t *recon_v, const vp8b200_segment_data *SD, float SSIM_target, int width,
                                       int height);

/* GPU half of count_probs + encode_coefficients (src/CPU_kernels.cl:347-778, enqueued src/vp8enc.cpp:65-88):
 * neighbour contexts (third_context [M][25]), per-partition token statistics coeff_probs /
 * coeff_probs_denom [P][4][8][3][11] exactly as count_probs leaves them, and for every partition the
 * stream of coded decisions in coding order (16-bit entries: bit 15 value, bits 0-10 probability slot
 * or 1056 + fixed probability).  part_info: [0,P) stream bases, [P,2P) counts, [2P] total; if the total
 * exceeds capacity no stream is written.  mb_tokens/mb_offset: int32 [M] scratch, tail_scratch: uint32 [P*68].
 * The serial bool coder over the streams runs on the host. */
int vp8b200_entropy_tokens(void *stream, const int16_t *MB, const int32_t *MB_non_zero_coeffs, const int32_t *MB_parts,
                           int mb_width, int mb_height, int num_partitions, uint32_t *coeff_probs,
                           uint32_t *coeff_probs_denom, uint8_t *third_context, uint16_t *tokens, uint32_t capacity,
                           int32_t *mb_tokens, int32_t *mb_offset, uint32_t *part_info, uint32_t *tail_scratch);

/* encode_coefficients (src/CPU_kernels.cl:541-778, enqueued src/vp8enc.cpp:88) over the decision streams of
 * vp8b200_entropy_tokens: RFC 6386's boolean coder in parallel (the range as a 128-state machine scanned over
 * chunks of decisions, the partition as one big sum, see entropy_kernels.cu).  Partition p is written at
 * output + p * partition_step, its byte count to partition_sizes[p] (0 if it does not fit its slot).
 * coeff_probs: the table the host wrote back ([4][8][3][11] uint32, first partition's slot).  max_decisions: an upper
 * bound of the decisions of any one partition (the total will do); scratch: device memory of
 * vp8b200_entropy_boolcode_scratch_bytes(max_decisions, num_partitions, partition_step) bytes.
 * Same bytes as the host bool coder; costs no host time. */
size_t vp8b200_entropy_boolcode_scratch_bytes(uint32_t max_decisions, int num_partitions, int partition_step);
int vp8b200_entropy_boolcode(void *stream, const uint16_t *tokens, const uint32_t *part_info,
                             const uint32_t *coeff_probs, uint8_t *output, int32_t *partition_sizes,
                             int num_partitions, int partition_step, uint32_t max_decisions, void *scratch);

/* replaces prepare_filter_mask, src/CPU_kernels.cl:782-827 (src/loop_filter.h:25-33) */
int vp8b200_prepare_filter_mask(void *stream, const int16_t *MB, int32_t *MB_non_zero_coeffs,
                                const int32_t *MB_parts, int32_t *mb_mask, int width, int height);

/* replaces loop_filter_frame_luma (mb_size 16) / loop_filter_frame_chroma (mb_size 8),
 * src/CPU_kernels.cl:829-1075,1333-1438 (src/loop_filter.h:140-183): the VP8 normal loop
 * filter in raster macroblock order, run as a wavefront over macroblock rows */
int vp8b200_loop_filter_frame(void *stream, uint8_t *frame, const int32_t *MB_segment_ids, const int32_t *mb_mask,
                              const vp8b200_segment_data *SD, int width, int height, int mb_size);

/* the three planes of one frame in a single launch (Y, U, V run concurrently) */
int vp8b200_loop_filter_planes(void *stream, uint8_t *y, uint8_t *u, uint8_t *v, const int32_t *MB_segment_ids,
                               const int32_t *mb_mask, const vp8b200_segment_data *SD, int width, int height);
/* The loop filter keeps a ticket counter and a mailbox per stream; whoever owns a stream releases them before
 * destroying it (vp8b200_engine_destroy does so for its own). */
void vp8b200_loop_filter_release(void *stream);

/* the intra (key-frame) path the reference host runs in plain C on one thread -- intra_transform() /
 * predict_and_transform_mb(), src/intra_part.h:37-741, 1089-1128 -- for a whole frame: every macroblock B_PRED (the
 * best of the ten 4x4 sub-block modes by the reference's weight()), chroma TM_PRED, forward DCT, quantise, dequantise,
 * inverse DCT, reconstruction, in the reference's raster dependency order (a wavefront over macroblocks on the GPU).
 * Planes are macroblock-padded (width, height multiples of 16).  Outputs as intra_transform() leaves them: rec_* the
 * unfiltered reconstruction, MB 25 x 16 int16 per macroblock in zig-zag order (block 24 untouched), modes 16 int32
 * per macroblock (frames.e_data[].mode), MB_parts = are4x4, MB_segment_id = intra_segment.  The four quantisers are
 * frames.y_dc_q / y_ac_q / uv_dc_q / uv_ac_q of the intra segment (src/vp8enc.cpp:160-181).  scratch: device memory
 * of vp8b200_intra_frame_scratch_bytes().  Not reachable from the unmodified host. */
size_t vp8b200_intra_frame_scratch_bytes(int width, int height);
int vp8b200_intra_frame(void *stream, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v, uint8_t *rec_y,
                        uint8_t *rec_u, uint8_t *rec_v, int16_t *MB, int32_t *modes, int32_t *MB_parts,
                        int32_t *MB_segment_id, int width, int height, int y_dc_q, int y_ac_q, int uv_dc_q, int uv_ac_q,
                        void *scratch);

/* the two O(N) reductions the reference host runs on every frame in plain C --
 * get_loopfilter_strength() (src/vp8enc.cpp:96-127: out4[0] = "reductor" from the mean luma, out4[1] = sharpness from
 * the mean squared difference of every interior pixel to the average of its eight neighbours) and the two chroma
 * differences scene_change() thresholds (src/vp8enc.cpp:265-285: out4[2] = mean |last_U - current_U|, out4[3] the same
 * for V, over the padded planes).  Same `int` arithmetic as the reference, wrap-around of its accumulators included.
 * cur_y may be NULL (chroma only) and so may the four chroma planes (luma only).  scratch4: four 64-bit words of
 * device memory, out4: four int32 of device memory.  Not reachable from the unmodified host. */
int vp8b200_frame_statistics(void *stream, const uint8_t *cur_y, int width, int height, const uint8_t *last_u,
                             const uint8_t *cur_u, const uint8_t *last_v, const uint8_t *cur_v,
                             unsigned long long *scratch4, int32_t *out4);

/* ------------------------------------------------------------------------------------------
 * Frame-level engine: one object owns every per-frame device buffer the reference creates in
 * init_all() (src/init.h:430-593: pyramids, LAST/GOLDEN/ALTREF planes, vector nets, metrics,
 * predictors, residuals, coefficient / vector / parts / reference / segment-id / SSIM arrays)
 * and runs the enqueue sequence of prepare_GPU_buffers() + inter_transform()
 * (src/inter_part.h:1-384) followed by the filter mask and the loop filter
 * (src/loop_filter.h:25-46,140-183) on one CUDA stream.  It is the same work the OpenCL shim
 * performs when the unmodified host drives it kernel by kernel.
 */
typedef struct vp8b200_engine vp8b200_engine;

/* stream: cudaStream_t as void*, NULL = a private stream */
vp8b200_engine *vp8b200_engine_create(int width, int height, void *stream);
void vp8b200_engine_destroy(vp8b200_engine *e);
void *vp8b200_engine_stream(vp8b200_engine *e);
int vp8b200_engine_synchronize(vp8b200_engine *e);

/* Seeds the reconstruction buffers (what the host uploads as LAST at src/vp8enc.cpp:395-401, e.g. the
 * loop-filtered key frame) from device (on_device=1) or host memory. */
int vp8b200_engine_set_reconstruction(vp8b200_engine *e, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                                      int on_device);

/* One inter frame with the current frame already resident in device memory.  On return (stream
 * order) the coefficient/vector/... buffers hold this frame's results and the reconstruction
 * buffers hold the UNFILTERED reconstruction.  prev_is_golden / prev_is_altref /
 * altref_differs_from_golden are the host's flags of src/inter_part.h:35-50,103-104. */
int vp8b200_engine_inter_frame(vp8b200_engine *e, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                               const vp8b200_segment_data *SD_host, float SSIM_target, int prev_is_golden,
                               int prev_is_altref, int altref_differs_from_golden);

/* One key frame with the current frame resident in device memory: intra_transform() (src/intra_part.h:1089-1128) on
 * the GPU (vp8b200_intra_frame) into the engine's coefficient, mode (VP8B200_BUF_INTRA_MODES), parts, segment-id and
 * reconstruction buffers; the quantisers are those of segment 0 of SD_host as prepare_segments_data() derives them on
 * the host (src/vp8enc.cpp:160-181).  Follow with vp8b200_engine_loop_filter(e, SD_host). */
int vp8b200_engine_key_frame(vp8b200_engine *e, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                             const vp8b200_segment_data *SD_host);

/* Filter mask + normal loop filter of the three reconstruction planes in place (they then are
 * the next frame's LAST). */
int vp8b200_engine_loop_filter(vp8b200_engine *e, const vp8b200_segment_data *SD_host);

/* The call with HOST buffers (pinned or pageable): uploads the current frame, runs
 * inter_frame + loop_filter, downloads every array the reference host reads back
 * (src/inter_part.h:263-265, src/vp8enc.cpp:422-433) plus the loop-filtered planes, and waits.
 * Any output pointer may be NULL. */
int vp8b200_engine_encode_frame_host(vp8b200_engine *e, const uint8_t *cur_y, const uint8_t *cur_u,
                                     const uint8_t *cur_v, const vp8b200_segment_data *SD, float SSIM_target,
                                     int prev_is_golden, int prev_is_altref, int altref_differs_from_golden,
                                     int16_t *MB_coeffs, int16_t *MB_vectors, int32_t *MB_parts,
                                     int32_t *MB_reference_frame, int32_t *MB_segment_id, float *MB_SSIM,
                                     int32_t *MB_non_zero_coeffs, uint8_t *recon_y, uint8_t *recon_u,
                                     uint8_t *recon_v);

/* device pointers of the engine's result buffers */
enum {
    VP8B200_BUF_COEFFS = 0, VP8B200_BUF_VECTORS, VP8B200_BUF_PARTS, VP8B200_BUF_REFERENCE_FRAME,
    VP8B200_BUF_SEGMENT_ID, VP8B200_BUF_SSIM, VP8B200_BUF_NON_ZERO, VP8B200_BUF_RECON_Y, VP8B200_BUF_RECON_U,
    VP8B200_BUF_RECON_V, VP8B200_BUF_INTRA_MODES, VP8B200_BUF_COUNT
};
void *vp8b200_engine_buffer(vp8b200_engine *e, int which);
/* number of kernels the last inter_frame + loop_filter launched */
int vp8b200_engine_last_launch_count(vp8b200_engine *e);

/* Measurement aid (bench.py's per-kernel roofline lines): with stage timing on, the engine records a CUDA event
 * on its stream at every stage boundary of a frame; vp8b200_engine_stage_times() waits for the stream and returns
 * the duration in milliseconds of every stage of the LAST frame (-1 for a stage that did not run).  The stages
 * follow the reference's enqueue sequence: buffer preparation (copies, reset_vectors, pyramid), the five
 * luma_search_1step levels, luma_search_2step, select_reference + pack, the predict / transform / SSIM ladder,
 * prepare_filter_mask, the loop filter. */
enum {
    VP8B200_STAGE_SETUP = 0, VP8B200_STAGE_SEARCH_16X, VP8B200_STAGE_SEARCH_8X, VP8B200_STAGE_SEARCH_4X,
    VP8B200_STAGE_SEARCH_2X, VP8B200_STAGE_SEARCH_1X, VP8B200_STAGE_SEARCH_QPEL, VP8B200_STAGE_SELECT,
    VP8B200_STAGE_TRANSFORM, VP8B200_STAGE_FILTER_MASK, VP8B200_STAGE_LOOP_FILTER, VP8B200_NUM_STAGES
};
int vp8b200_engine_stage_timing(vp8b200_engine *e, int on);
int vp8b200_engine_stage_times(vp8b200_engine *e, float *ms, int cap);

#ifdef __cplusplus
}
#endif
#endif /* VP8B200_H */
