// Frame-level engine (see include/vp8b200.h): owns the per-frame device buffers and runs the
// reference's enqueue sequence (src/inter_part.h:1-384, src/loop_filter.h) on one CUDA stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

struct vp8b200_engine {
    int w, h, M;
    cudaStream_t stream;
    bool own_stream;
    // index k: plane down-sampled by 2^k.  last_pyr[0] is reconstructed_frame_Y.
    uint8_t *cur_pyr[5], *last_pyr[5], *gold_pyr[5], *alt_pyr[5];
    uint8_t *recon_u, *recon_v;
    uint8_t *cur_u_stage, *cur_v_stage;  // current chroma of the host-buffer entry point
    uint8_t *img[3][3];  // [LAST/GOLDEN/ALTREF][Y/U/V]: the reference's image objects
    int16_t *net[3][2];
    int32_t *metrics[3];
    uint8_t *pred[3];
    int16_t *res[3];
    int16_t *coeffs, *vectors;
    int32_t *parts, *ref_frame, *seg_id, *nz, *mask;
    int32_t *modes;       // key frames: the sixteen sub-block modes of every macroblock
    void *intra_scratch;  // row progress counters of the intra kernel
    float *ssim;
    vp8b200_segment_data *sd_dev;
    vp8b200_segment_data *sd_pinned;  // 8 slots of 4 segments
    cudaEvent_t sd_event[8];
    int sd_next;
    int launches;
    bool fused;  // VP8B200_FUSED=0 selects the kernel-per-kernel sequence
    // stage timing (vp8b200_engine_stage_timing): events at the stage boundaries of a frame on the engine's stream
    bool timing;
    cudaEvent_t stage_ev[VP8B200_NUM_STAGES + 1];
    bool stage_seen[VP8B200_NUM_STAGES + 1];
};

namespace {
template <class T>
bool dalloc(T *&p, size_t bytes) {
    if (cudaMalloc((void **)&p, bytes ? bytes : 1) != cudaSuccess) return false;
    return cudaMemset(p, 0, bytes ? bytes : 1) == cudaSuccess;
}
#define TRY(x)            \
    do {                  \
        int rc_ = (x);    \
        if (rc_) return rc_; \
        ++e->launches;    \
    } while (0)
inline int cu(cudaError_t err) { return err == cudaSuccess ? 0 : -(int)err; }
// start of stage i (= end of stage i - 1)
inline void stage(vp8b200_engine *e, int i) {
    if (!e->timing) return;
    cudaEventRecord(e->stage_ev[i], e->stream);
    e->stage_seen[i] = true;
}
}  // namespace

extern "C" int vp8b200_engine_stage_timing(vp8b200_engine *e, int on) {
    if (!e) return -1;
    if (on && !e->stage_ev[0])
        for (int i = 0; i <= VP8B200_NUM_STAGES; ++i)
            if (cudaEventCreate(&e->stage_ev[i]) != cudaSuccess) return -(int)cudaGetLastError();
    e->timing = on != 0;
    for (int i = 0; i <= VP8B200_NUM_STAGES; ++i) e->stage_seen[i] = false;
    return 0;
}

extern "C" int vp8b200_engine_stage_times(vp8b200_engine *e, float *ms, int cap) {
    if (!e || !e->timing) return -1;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) return -(int)cudaGetLastError();
    for (int i = 0; i < VP8B200_NUM_STAGES && i < cap; ++i) {
        ms[i] = -1.0f;
        if (e->stage_seen[i] && e->stage_seen[i + 1]) cudaEventElapsedTime(&ms[i], e->stage_ev[i], e->stage_ev[i + 1]);
    }
    return 0;
}

extern "C" vp8b200_engine *vp8b200_engine_create(int width, int height, void *stream) {
    if (width < 16 || height < 16 || (width & 15) || (height & 15)) return nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return nullptr;  // no fallback
    vp8b200_engine *e = new vp8b200_engine();
    memset(e, 0, sizeof(*e));
    e->w = width;
    e->h = height;
    e->M = (width / 16) * (height / 16);
    { const char *f = getenv("VP8B200_FUSED"); e->fused = !(f && f[0] == '0'); }
    if (stream) {
        e->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete e;
            return nullptr;
        }
        e->own_stream = true;
    }
    const size_t ysz = (size_t)width * height, csz = ysz / 4;
    bool ok = true;
    for (int k = 0; k < 5; ++k) {
        const size_t sz = (size_t)(width >> k) * (height >> k);
        ok = ok && dalloc(e->cur_pyr[k], sz) && dalloc(e->last_pyr[k], sz) && dalloc(e->gold_pyr[k], sz) &&
             dalloc(e->alt_pyr[k], sz);
    }
    ok = ok && dalloc(e->recon_u, csz) && dalloc(e->recon_v, csz) && dalloc(e->cur_u_stage, csz) && dalloc(e->cur_v_stage, csz);
    for (int r = 0; r < 3; ++r) {
        for (int p = 0; p < 3; ++p) ok = ok && dalloc(e->img[r][p], p ? csz : ysz);
        for (int k = 0; k < 2; ++k) ok = ok && dalloc(e->net[r][k], (size_t)e->M * 16);
        ok = ok && dalloc(e->metrics[r], (size_t)e->M * 16);
    }
    for (int p = 0; p < 3; ++p) ok = ok && dalloc(e->pred[p], p ? csz : ysz) && dalloc(e->res[p], (p ? csz : ysz) * 2);
    ok = ok && dalloc(e->coeffs, (size_t)e->M * 800) && dalloc(e->vectors, (size_t)e->M * 16) &&
         dalloc(e->parts, (size_t)e->M * 4) && dalloc(e->ref_frame, (size_t)e->M * 4) &&
         dalloc(e->seg_id, (size_t)e->M * 4) && dalloc(e->nz, (size_t)e->M * 4) && dalloc(e->mask, (size_t)e->M * 4) &&
         dalloc(e->modes, (size_t)e->M * 64) && dalloc(e->intra_scratch, vp8b200_intra_frame_scratch_bytes(e->w, e->h)) &&
         dalloc(e->ssim, (size_t)e->M * 4) && dalloc(e->sd_dev, sizeof(vp8b200_segment_data) * 4);
    ok = ok && cudaHostAlloc((void **)&e->sd_pinned, sizeof(vp8b200_segment_data) * 4 * 8, cudaHostAllocDefault) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreateWithFlags(&e->sd_event[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        vp8b200_engine_destroy(e);
        return nullptr;
    }
    return e;
}

extern "C" void vp8b200_engine_destroy(vp8b200_engine *e) {
    if (!e) return;
    cudaStreamSynchronize(e->stream);
    vp8b200_loop_filter_release(e->stream);
    for (int k = 0; k < 5; ++k) {
        cudaFree(e->cur_pyr[k]); cudaFree(e->last_pyr[k]); cudaFree(e->gold_pyr[k]); cudaFree(e->alt_pyr[k]);
    }
    cudaFree(e->recon_u); cudaFree(e->recon_v); cudaFree(e->cur_u_stage); cudaFree(e->cur_v_stage);
    for (int r = 0; r < 3; ++r) {
        for (int p = 0; p < 3; ++p) cudaFree(e->img[r][p]);
        for (int k = 0; k < 2; ++k) cudaFree(e->net[r][k]);
        cudaFree(e->metrics[r]);
    }
    for (int p = 0; p < 3; ++p) { cudaFree(e->pred[p]); cudaFree(e->res[p]); }
    cudaFree(e->coeffs); cudaFree(e->vectors); cudaFree(e->parts); cudaFree(e->ref_frame); cudaFree(e->seg_id);
    cudaFree(e->nz); cudaFree(e->mask); cudaFree(e->ssim); cudaFree(e->sd_dev); cudaFree(e->modes); cudaFree(e->intra_scratch);
    for (int i = 0; i <= VP8B200_NUM_STAGES; ++i)
        if (e->stage_ev[i]) cudaEventDestroy(e->stage_ev[i]);
    if (e->sd_pinned) cudaFreeHost(e->sd_pinned);
    for (int i = 0; i < 8; ++i)
        if (e->sd_event[i]) cudaEventDestroy(e->sd_event[i]);
    if (e->own_stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" void *vp8b200_engine_stream(vp8b200_engine *e) { return e ? (void *)e->stream : nullptr; }
extern "C" int vp8b200_engine_synchronize(vp8b200_engine *e) { return cu(cudaStreamSynchronize(e->stream)); }
extern "C" int vp8b200_engine_last_launch_count(vp8b200_engine *e) { return e->launches; }

extern "C" void *vp8b200_engine_buffer(vp8b200_engine *e, int which) {
    switch (which) {
        case VP8B200_BUF_COEFFS: return e->coeffs;
        case VP8B200_BUF_VECTORS: return e->vectors;
        case VP8B200_BUF_PARTS: return e->parts;
        case VP8B200_BUF_REFERENCE_FRAME: return e->ref_frame;
        case VP8B200_BUF_SEGMENT_ID: return e->seg_id;
        case VP8B200_BUF_SSIM: return e->ssim;
        case VP8B200_BUF_NON_ZERO: return e->nz;
        case VP8B200_BUF_RECON_Y: return e->last_pyr[0];
        case VP8B200_BUF_RECON_U: return e->recon_u;
        case VP8B200_BUF_RECON_V: return e->recon_v;
        case VP8B200_BUF_INTRA_MODES: return e->modes;
        default: return nullptr;
    }
}

extern "C" int vp8b200_engine_set_reconstruction(vp8b200_engine *e, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                                                 int on_device) {
    const size_t ysz = (size_t)e->w * e->h, csz = ysz / 4;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int rc = cu(cudaMemcpyAsync(e->last_pyr[0], y, ysz, kind, e->stream));
    if (!rc) rc = cu(cudaMemcpyAsync(e->recon_u, u, csz, kind, e->stream));
    if (!rc) rc = cu(cudaMemcpyAsync(e->recon_v, v, csz, kind, e->stream));
    return rc;
}

static int upload_sd(vp8b200_engine *e, const vp8b200_segment_data *SD_host, int) {
    // through a ring of pinned staging slots, so that the copy is asynchronous, the caller may reuse
    // SD_host at once and frames can be queued without a host-device round trip
    const int slot = e->sd_next;
    e->sd_next = (e->sd_next + 1) % 8;
    cudaEventSynchronize(e->sd_event[slot]);  // only waits if the ring has been lapped
    memcpy(e->sd_pinned + 4 * slot, SD_host, sizeof(vp8b200_segment_data) * 4);
    int rc = cu(cudaMemcpyAsync(e->sd_dev, e->sd_pinned + 4 * slot, sizeof(vp8b200_segment_data) * 4,
                                cudaMemcpyHostToDevice, e->stream));
    if (!rc) rc = cu(cudaEventRecord(e->sd_event[slot], e->stream));
    return rc;
}

static int inter_frame_body(vp8b200_engine *e, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                            float SSIM_target, int prev_is_golden, int prev_is_altref, int altref_differs) {
    const int w = e->w, h = e->h, M = e->M;
    const size_t ysz = (size_t)w * h, csz = ysz / 4;
    void *s = e->stream;
    const int use_golden = !prev_is_golden;
    const int use_altref = !prev_is_altref && altref_differs;  // src/inter_part.h:103-104
    const uint8_t *cur[3] = {cur_y, cur_u, cur_v};
    uint8_t *recon[3] = {e->last_pyr[0], e->recon_u, e->recon_v};
    e->launches = 0;
    stage(e, VP8B200_STAGE_SETUP);

    // the host's uploads of the previous (loop-filtered) reconstruction into the image objects
    // (src/vp8enc.cpp:399-401) are device-to-device copies here
    for (int p = 0; p < 3; ++p)
        if (cudaMemcpyAsync(e->img[vp8::LAST][p], recon[p], p ? csz : ysz, cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess)
            return -1;

    // prepare_GPU_buffers(), src/inter_part.h:1-94
    TRY(vp8b200_reset_vectors(s, e->net[0][0], e->net[0][1], e->net[1][0], e->net[1][1], e->net[2][0], e->net[2][1],
                              e->metrics[0], e->metrics[1], e->metrics[2], M * 4));
    const uint8_t *cpyr[5] = {cur_y, e->cur_pyr[1], e->cur_pyr[2], e->cur_pyr[3], e->cur_pyr[4]};
    for (int k = 0; k < 4; ++k) {
        TRY(vp8b200_downsample_x2(s, e->last_pyr[k], e->last_pyr[k + 1], w >> k, h >> k));
        TRY(vp8b200_downsample_x2(s, cpyr[k], e->cur_pyr[k + 1], w >> k, h >> k));
    }
    if (prev_is_golden) {
        for (int k = 0; k < 5; ++k)
            cudaMemcpyAsync(e->gold_pyr[k], e->last_pyr[k], (size_t)(w >> k) * (h >> k), cudaMemcpyDeviceToDevice, e->stream);
        for (int p = 0; p < 3; ++p)
            cudaMemcpyAsync(e->img[vp8::GOLDEN][p], e->img[vp8::LAST][p], p ? csz : ysz, cudaMemcpyDeviceToDevice, e->stream);
    }
    if (prev_is_altref) {
        for (int k = 0; k < 5; ++k)
            cudaMemcpyAsync(e->alt_pyr[k], e->last_pyr[k], (size_t)(w >> k) * (h >> k), cudaMemcpyDeviceToDevice, e->stream);
        for (int p = 0; p < 3; ++p)
            cudaMemcpyAsync(e->img[vp8::ALTREF][p], e->img[vp8::LAST][p], p ? csz : ysz, cudaMemcpyDeviceToDevice, e->stream);
    }

    // pyramid search, src/inter_part.h:109-236; nets ping-pong 1->2, 2->1, 1->2, 2->1, 1->2, qpel 2->1 (Q4)
    uint8_t **pyr[3] = {e->last_pyr, e->gold_pyr, e->alt_pyr};
    const int use[3] = {1, use_golden, use_altref};
    if (e->fused) {
        // the references in use are searched by one launch per level (grid.y = reference)
        int nrefs = 0;
        const uint8_t *rp[3];
        const int16_t *sn[3];
        int16_t *dn[3];
        int32_t *met[3];
        int rid[3];
        for (int r = 0; r < 3; ++r)
            if (use[r]) rid[nrefs++] = r;
        for (int k = 4; k >= 0; --k) {
            const int src = (k & 1) ? 1 : 0;
            for (int i = 0; i < nrefs; ++i) {
                rp[i] = pyr[rid[i]][k];
                sn[i] = e->net[rid[i]][src];
                dn[i] = e->net[rid[i]][src ^ 1];
            }
            stage(e, VP8B200_STAGE_SEARCH_16X + (4 - k));
            TRY(vp8b200_luma_search_1step_multi(s, cpyr[k], nrefs, rp, sn, dn, (w / 16) * 2, w >> k, h >> k, 1 << k));
        }
        for (int i = 0; i < nrefs; ++i) {
            rp[i] = e->img[rid[i]][0];
            sn[i] = e->net[rid[i]][1];
            dn[i] = e->net[rid[i]][0];
            met[i] = e->metrics[rid[i]];
        }
        stage(e, VP8B200_STAGE_SEARCH_QPEL);
        TRY(vp8b200_luma_search_2step_multi(s, cur_y, nrefs, rp, sn, dn, met, w, h));
    } else {
        for (int k = 4; k >= 0; --k) {
            const int src = (k & 1) ? 1 : 0;
            stage(e, VP8B200_STAGE_SEARCH_16X + (4 - k));
            for (int r = 0; r < 3; ++r)
                if (use[r])
                    TRY(vp8b200_luma_search_1step(s, cpyr[k], pyr[r][k], e->net[r][src], e->net[r][src ^ 1], (w / 16) * 2,
                                                  w >> k, h >> k, 1 << k));
        }
        stage(e, VP8B200_STAGE_SEARCH_QPEL);
        for (int r = 0; r < 3; ++r)
            if (use[r])
                TRY(vp8b200_luma_search_2step(s, cur_y, e->img[r][0], e->net[r][1], e->net[r][0], e->metrics[r], w, h));
    }

    stage(e, VP8B200_STAGE_SELECT);
    TRY(vp8b200_select_reference(s, e->net[0][0], e->net[1][0], e->net[2][0], e->metrics[0], e->metrics[1], e->metrics[2],
                                 e->ref_frame, e->vectors, w, h, use_golden, use_altref));
    TRY(vp8b200_pack_8x8_into_16x16(s, e->vectors, e->parts, e->ssim, M));
    stage(e, VP8B200_STAGE_TRANSFORM);

    if (e->fused && SSIM_target >= -2.0f) {
        const uint8_t *img[9];
        for (int r = 0; r < 3; ++r)
            for (int p = 0; p < 3; ++p) img[3 * r + p] = use[r] ? e->img[r][p] : nullptr;
        TRY(vp8b200_mb_predict_transform_fused(s, cur_y, cur_u, cur_v, img, e->ref_frame, e->vectors, e->parts, e->coeffs,
                                               e->seg_id, e->ssim, recon[0], recon[1], recon[2], e->sd_dev, SSIM_target,
                                               w, h));
        stage(e, VP8B200_STAGE_FILTER_MASK);
        return 0;
    }
    for (int p = 0; p < 3; ++p)
        for (int r = 0; r < 3; ++r)
            if (use[r])
                TRY(vp8b200_prepare_predictors_and_residual(s, cur[p], e->img[r][p], e->pred[p], e->res[p], e->ref_frame,
                                                            e->vectors, p ? w / 2 : w, p ? h / 2 : h, p, r));

    // the SSIM-driven re-quantisation ladder, src/inter_part.h:329-378 (Q10)
    for (int seg = 3; seg >= 0; --seg) {
        for (int p = 0; p < 3; ++p)
            TRY(vp8b200_dct4x4(s, e->res[p], e->coeffs, e->seg_id, e->parts, e->ssim, p ? w / 2 : w, p ? h / 2 : h,
                               e->sd_dev, seg, SSIM_target, p));
        TRY(vp8b200_wht4x4_iwht4x4(s, e->coeffs, e->seg_id, e->parts, e->sd_dev, seg, M));
        for (int p = 0; p < 3; ++p)
            TRY(vp8b200_idct4x4(s, recon[p], e->pred[p], e->coeffs, e->seg_id, e->parts, p ? w / 2 : w, p ? h / 2 : h,
                                e->sd_dev, seg, p));
        for (int p = 0; p < 3; ++p)
            TRY(vp8b200_count_SSIM(s, cur[p], recon[p], e->seg_id, (float *)e->metrics[p], p ? w / 2 : w, p ? h / 2 : h,
                                   seg, p ? 8 : 16));
        TRY(vp8b200_gather_SSIM(s, (float *)e->metrics[0], (float *)e->metrics[1], (float *)e->metrics[2], e->ssim, M));
    }
    stage(e, VP8B200_STAGE_FILTER_MASK);
    return 0;
}

extern "C" int vp8b200_engine_inter_frame(vp8b200_engine *e, const uint8_t *cur_y, const uint8_t *cur_u,
                                          const uint8_t *cur_v, const vp8b200_segment_data *SD_host, float SSIM_target,
                                          int prev_is_golden, int prev_is_altref, int altref_differs) {
    int rc = upload_sd(e, SD_host, 0);
    if (rc) return rc;
    return inter_frame_body(e, cur_y, cur_u, cur_v, SSIM_target, prev_is_golden, prev_is_altref, altref_differs);
}

// A key frame (intra_transform(), src/intra_part.h:1089-1128, as the host runs it for frame 0, at every GOP boundary and
// for forced keys): the current frame is coded intra into the engine's coefficient / reconstruction buffers, with the
// quantisers of segment 0 derived from SD_host the way prepare_segments_data() derives frames.y_dc_q ...
// (src/vp8enc.cpp:160-181: the host-side table, where uv_dc is capped at 132 but y2 plays no role).  The loop filter
// (vp8b200_engine_loop_filter with the same SD) then makes the reconstruction the LAST reference; GOLDEN and ALTREF
// follow on the next inter frame through its prev_is_golden / prev_is_altref flags, as in the reference.
extern "C" int vp8b200_engine_key_frame(vp8b200_engine *e, const uint8_t *cur_y, const uint8_t *cur_u, const uint8_t *cur_v,
                                        const vp8b200_segment_data *SD_host) {
    if (!e || !SD_host) return -(int)cudaErrorInvalidValue;
    int rc = upload_sd(e, SD_host, 0);
    if (rc) return rc;
    static const short dcq[128] = {
        4,   5,   6,   7,   8,   9,   10,  10,  11,  12,  13,  14,  15,  16,  17,  17,  18,  19,  20,  20,  21,  21,
        22,  22,  23,  23,  24,  25,  25,  26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  37,  38,  39,
        40,  41,  42,  43,  44,  45,  46,  46,  47,  48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  59,  60,
        61,  62,  63,  64,  65,  66,  67,  68,  69,  70,  71,  72,  73,  74,  75,  76,  76,  77,  78,  79,  80,  81,
        82,  83,  84,  85,  86,  87,  88,  89,  91,  93,  95,  96,  98,  100, 101, 102, 104, 106, 108, 110, 112, 114,
        116, 118, 122, 124, 126, 128, 130, 132, 134, 136, 138, 140, 143, 145, 148, 151, 154, 157};
    static const short acq[128] = {
        4,   5,   6,   7,   8,   9,   10,  11,  12,  13,  14,  15,  16,  17,  18,  19,  20,  21,  22,  23,  24,  25,
        26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  38,  39,  40,  41,  42,  43,  44,  45,  46,  47,
        48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  60,  62,  64,  66,  68,  70,  72,  74,  76,  78,  80,
        82,  84,  86,  88,  90,  92,  94,  96,  98,  100, 102, 104, 106, 108, 110, 112, 114, 116, 119, 122, 125, 128,
        131, 134, 137, 140, 143, 146, 149, 152, 155, 158, 161, 164, 167, 170, 173, 177, 181, 185, 189, 193, 197, 201,
        205, 209, 213, 217, 221, 225, 229, 234, 239, 245, 249, 254, 259, 264, 269, 274, 279, 284};
    auto cl = [](int v) { return v < 0 ? 0 : (v > 127 ? 127 : v); };
    const int i = SD_host[0].y_ac_i;
    const int y_ac = acq[cl(i)], y_dc = dcq[cl(i + SD_host[0].y_dc_idelta)];
    int uv_dc = dcq[cl(i + SD_host[0].uv_dc_idelta)];
    const int uv_ac = acq[cl(i + SD_host[0].uv_ac_idelta)];
    if (uv_dc > 132) uv_dc = 132;
    e->launches = 0;
    TRY(vp8b200_intra_frame(e->stream, cur_y, cur_u, cur_v, e->last_pyr[0], e->recon_u, e->recon_v, e->coeffs, e->modes, e->parts,
                            e->seg_id, e->w, e->h, y_dc, y_ac, uv_dc, uv_ac, e->intra_scratch));
    return 0;
}

extern "C" int vp8b200_engine_loop_filter(vp8b200_engine *e, const vp8b200_segment_data *SD_host) {
    if (SD_host) {
        int rc = upload_sd(e, SD_host, 1);
        if (rc) return rc;
    }
    e->launches = 0;
    stage(e, VP8B200_STAGE_FILTER_MASK);
    TRY(vp8b200_prepare_filter_mask(e->stream, e->coeffs, e->nz, e->parts, e->mask, e->w, e->h));
    stage(e, VP8B200_STAGE_LOOP_FILTER);
    TRY(vp8b200_loop_filter_planes(e->stream, e->last_pyr[0], e->recon_u, e->recon_v, e->seg_id, e->mask, e->sd_dev, e->w,
                                   e->h));
    stage(e, VP8B200_NUM_STAGES);
    return 0;
}

extern "C" int vp8b200_engine_encode_frame_host(vp8b200_engine *e, const uint8_t *cur_y, const uint8_t *cur_u,
                                                const uint8_t *cur_v, const vp8b200_segment_data *SD, float SSIM_target,
                                                int prev_is_golden, int prev_is_altref, int altref_differs,
                                                int16_t *MB_coeffs, int16_t *MB_vectors, int32_t *MB_parts,
                                                int32_t *MB_reference_frame, int32_t *MB_segment_id, float *MB_SSIM,
                                                int32_t *MB_non_zero_coeffs, uint8_t *recon_y, uint8_t *recon_u,
                                                uint8_t *recon_v) {
    const size_t ysz = (size_t)e->w * e->h, csz = ysz / 4, M = (size_t)e->M;
    cudaStream_t st = e->stream;
    int rc = upload_sd(e, SD, 0);
    if (!rc) rc = cu(cudaMemcpyAsync(e->cur_pyr[0], cur_y, ysz, cudaMemcpyHostToDevice, st));
    if (!rc) rc = cu(cudaMemcpyAsync(e->cur_u_stage, cur_u, csz, cudaMemcpyHostToDevice, st));
    if (!rc) rc = cu(cudaMemcpyAsync(e->cur_v_stage, cur_v, csz, cudaMemcpyHostToDevice, st));
    if (!rc) rc = inter_frame_body(e, e->cur_pyr[0], e->cur_u_stage, e->cur_v_stage, SSIM_target, prev_is_golden,
                                   prev_is_altref, altref_differs);
    if (rc) return rc;
    const int n_inter = e->launches;
#define D2H(dst, src, bytes) \
    if ((dst) && !rc) rc = cu(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, st))
    D2H(MB_vectors, e->vectors, M * 16);
    D2H(MB_parts, e->parts, M * 4);
    D2H(MB_reference_frame, e->ref_frame, M * 4);
    D2H(MB_coeffs, e->coeffs, M * 800);
    D2H(MB_segment_id, e->seg_id, M * 4);
    D2H(MB_SSIM, e->ssim, M * 4);
    if (rc) return rc;
    rc = vp8b200_engine_loop_filter(e, nullptr);  // same segment data as the transform
    if (rc) return rc;
    e->launches += n_inter;
    D2H(MB_non_zero_coeffs, e->nz, M * 4);
    D2H(recon_y, e->last_pyr[0], ysz);
    D2H(recon_u, e->recon_u, csz);
    D2H(recon_v, e->recon_v, csz);
#undef D2H
    if (rc) return rc;
    return cu(cudaStreamSynchronize(st));
}
