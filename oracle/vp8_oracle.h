/*
 * oracle/vp8_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the device kernels on vp8oclenc's inter-frame hot path
 * (SURVEY.md section 8a).  One function per reference __kernel, same argument
 * meaning, whole-NDRange semantics (the function does what enqueueing the kernel
 * over its full global size does).  Each function cites the reference lines it
 * follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may use this library.
 *
 * PINNING: the reference ships no tests or golden vectors.
 * This restatement is pinned against the reference ITSELF: oracle/_ref holds the
 * reference's own GPU_kernels.cl / CPU_kernels.cl compiled for the host CPU
 * (oracle/Makefile, target "ref"); tests/test_oracle_vs_ref.py runs every function
 * below against the corresponding reference kernel on seeded inputs, and
 * tests/golden/ holds vectors generated from the reference (tools/make_golden.py).
 */
#ifndef VP8_ORACLE_H
#define VP8_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/vp8enc.h:80-92 == src/GPU_kernels.cl:24-36 */
typedef struct {
    int32_t y_ac_i;
    int32_t y_dc_idelta;
    int32_t y2_dc_idelta;
    int32_t y2_ac_idelta;
    int32_t uv_dc_idelta;
    int32_t uv_ac_idelta;
    int32_t loop_filter_level;
    int32_t mbedge_limit;
    int32_t sub_bedge_limit;
    int32_t interior_limit;
    int32_t hev_threshold;
} vp8o_segment_data;

enum { VP8O_ARE16x16 = 0, VP8O_ARE8x8 = 1, VP8O_ARE4x4 = 2 };
enum { VP8O_LAST = 0, VP8O_GOLDEN = 1, VP8O_ALTREF = 2 };

/* src/GPU_kernels.cl:85-190 (with the s4..s7 clobber).  r = 4x4 residual, row-major */
int vp8o_weight(const int r[16]);

/* src/GPU_kernels.cl:404-427; n = number of 8x8 net entries (4 per MB) */
void vp8o_reset_vectors(int16_t *last1, int16_t *last2, int16_t *gold1, int16_t *gold2, int16_t *alt1,
                        int16_t *alt2, int32_t *last_Bdiff, int32_t *gold_Bdiff, int32_t *alt_Bdiff, int n);

/* src/GPU_kernels.cl:429-451 */
void vp8o_downsample_x2(const uint8_t *src, uint8_t *dst, int src_width, int src_height);

/* src/GPU_kernels.cl:459-560.  nets are short2 arrays (x,y interleaved) */
void vp8o_luma_search_1step(const uint8_t *cur, const uint8_t *prev, const int16_t *src_net, int16_t *dst_net,
                            int net_width, int width, int height, int pixel_rate);

/* src/GPU_kernels.cl:776-1203 (construct_opt1/2 + luma_search_2step); ref is the W x H image */
void vp8o_luma_search_2step(const uint8_t *cur, const uint8_t *ref, const int16_t *net, int16_t *ref_net,
                            int32_t *ref_Bdiff, int width, int height);

/* src/GPU_kernels.cl:1205-1283.  MB_vectors: 4 short2 per MB */
void vp8o_select_reference(const int16_t *last_net, const int16_t *gold_net, const int16_t *alt_net,
                           const int32_t *last_Bdiff, const int32_t *gold_Bdiff, const int32_t *alt_Bdiff,
                           int32_t *MB_reference_frame, int16_t *MB_vectors, int width, int height,
                           int use_golden, int use_altref);

/* src/GPU_kernels.cl:1346-1366 */
void vp8o_pack_8x8_into_16x16(const int16_t *MB_vectors, int32_t *MB_parts, float *MB_SSIM, int mb_count);

/* src/GPU_kernels.cl:574-774 (construct) + 1285-1344.  width/height are the PLANE's */
void vp8o_prepare_predictors_and_residual(const uint8_t *cur, const uint8_t *ref, uint8_t *predictor,
                                          int16_t *residual, const int32_t *MB_reference_frame,
                                          const int16_t *MB_vectors, int width, int height, int plane, int ref_id);

/* src/GPU_kernels.cl:1368-1496.  MB: 400 int16 per macroblock (25 blocks x 16) */
void vp8o_dct4x4(const int16_t *residual, int16_t *MB, int32_t *MB_segment_id, const int32_t *MB_parts,
                 const float *MB_SSIM, int width, int height, const vp8o_segment_data *SD, int segment_id,
                 float SSIM_target, int plane);

/* src/GPU_kernels.cl:257-401 + 1498-1543 */
void vp8o_wht4x4_iwht4x4(int16_t *MB, int32_t *MB_segment_id, const int32_t *MB_parts,
                         const vp8o_segment_data *SD, int segment_id, int mb_count);

/* src/GPU_kernels.cl:192-255 + 1545-1608 */
void vp8o_idct4x4(uint8_t *recon, const uint8_t *predictor, const int16_t *MB, const int32_t *MB_segment_id,
                  const int32_t *MB_parts, int width, int height, const vp8o_segment_data *SD, int segment_id,
                  int plane);

/* src/GPU_kernels.cl:1610-1971 (mb_size 16) and 1973-2095 (mb_size 8) */
void vp8o_count_SSIM(const uint8_t *frame1, const uint8_t *frame2, const int32_t *MB_segment_id, float *metric,
                     int width, int height, int segment_id, int mb_size);

/* src/GPU_kernels.cl:2097-2105 */
void vp8o_gather_SSIM(const float *m1, const float *m2, const float *m3, float *MB_SSIM, int mb_count);

/* src/CPU_kernels.cl:782-827 */
void vp8o_prepare_filter_mask(const int16_t *MB, int32_t *MB_non_zero_coeffs, const int32_t *MB_parts,
                              int32_t *mb_mask, int width, int height);

/* src/CPU_kernels.cl:829-1075 (mb_size 16) and 1333-1438 (mb_size 8) */
void vp8o_loop_filter_frame(uint8_t *frame, const int32_t *MB_segment_ids, const int32_t *mb_mask,
                            const vp8o_segment_data *SD, int width, int height, int mb_size);

/* ---- whole inter frame, in the enqueue order of src/inter_part.h:1-384 ------------------ */
typedef struct vp8o_frame_ctx vp8o_frame_ctx;
vp8o_frame_ctx *vp8o_ctx_create(int width, int height);
void vp8o_ctx_destroy(vp8o_frame_ctx *c);
/* Encodes one inter frame.  recon_* hold the previous (loop-filtered) reconstruction on entry
 * (what the host uploads at src/vp8enc.cpp:395-401) and the new unfiltered reconstruction on
 * exit (what it reads back at :431-433).  prev_is_golden / prev_is_altref / golden!=altref are
 * the host flags of src/inter_part.h:35-50,103-104. */
void vp8o_inter_frame(vp8o_frame_ctx *c, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                      uint8_t *recon_y, uint8_t *recon_u, uint8_t *recon_v, const vp8o_segment_data *SD,
                      float SSIM_target, int prev_is_golden, int prev_is_altref, int altref_differs_from_golden,
                      int16_t *MB_coeffs, int16_t *MB_vectors, int32_t *MB_parts, int32_t *MB_reference_frame,
                      int32_t *MB_segment_id, float *MB_SSIM);
/* mask + normal loop filter of the three planes (src/loop_filter.h:25-46,140-183) */
void vp8o_loop_filter_planes(uint8_t *y, uint8_t *u, uint8_t *v, const int16_t *MB_coeffs,
                             const int32_t *MB_parts, const int32_t *MB_segment_id, const vp8o_segment_data *SD,
                             int32_t *MB_non_zero_coeffs, int width, int height);

/* src/vp8enc.cpp:96-127 (get_loopfilter_strength) and :265-285 (the differences scene_change() thresholds);
 * out4 = { reductor, sharpness, Udiff, Vdiff }; either the luma plane or the chroma planes may be NULL */
void vp8o_frame_statistics(const uint8_t *cur_y, int width, int height, const uint8_t *last_u, const uint8_t *cur_u,
                           const uint8_t *last_v, const uint8_t *cur_v, int32_t *out4);

/* the intra (key-frame) path of the reference host, src/intra_part.h:37-741, 1089-1128 (vp8_oracle_intra.c):
 * quants = { y_dc_q, y_ac_q, uv_dc_q, uv_ac_q }; MB 25 x 16 int16 per macroblock (zig-zag), modes 16 per macroblock */
void vp8o_intra_frame(int width, int height, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                      uint8_t *rec_y, uint8_t *rec_u, uint8_t *rec_v, int16_t *MB, int32_t *modes, int32_t *parts,
                      int32_t *segment_id, const int32_t *quants);

/* test aid (tests/test_decoder_pin.py): which = 0 -> blocks whose predictor Q5 changed (x, y, plane), which = 1 ->
 * loop-filter edges that chained an unclamped value, Q7 (x, y, macroblock size).  Returns the number of events
 * since the last reset; at most `cap` positions are copied. */
void vp8o_quirk_log_reset(void);
int vp8o_quirk_log_get(int which, int *xyt, int cap);

#ifdef __cplusplus
}
#endif
#endif
