// TEST INFRASTRUCTURE -- the REFERENCE's own intra path, compiled from where it lies.
//
// intra_transform() / predict_and_transform_mb() (src/intra_part.h:517-741, 1089-1128) are plain C inside the
// reference's host translation unit and work on its global `frames` / `video` structures.  This file is that
// translation unit's skeleton: the same globals, the reference's headers included unmodified ($(REF)/vp8enc.h,
// $(REF)/intra_part.h), and one entry point that fills the globals from caller buffers and runs the reference's
// function over a whole frame.  It pins oracle/vp8_oracle.c's vp8o_intra_frame (tests/test_oracle_vs_ref.py).
// Built by `make -C oracle ref` into oracle/_ref/libref_intra.so; nothing here is copied from the reference.
#include <cstdlib>
#include <cstring>

#include "vp8enc.h"

struct fileContext input_file, output_file, error_file, dump_file;
struct deviceContext device;
struct videoContext video;
struct hostFrameBuffers frames;
struct encoderStatistics encStat;
static cl_int ifFlush(cl_command_queue) { return 0; }
static cl_int finalFlush(cl_command_queue) { return 0; }

#include "intra_part.h"

// quants = { y_dc_q, y_ac_q, uv_dc_q, uv_ac_q } of the intra segment.  Planes are wrk-sized (multiples of 16).
// MB_out: 25 blocks x 16 coefficients (zig-zag order) per macroblock, modes_out: 16 sub-block modes per macroblock.
extern "C" void vp8ref_intra_frame(int width, int height, const unsigned char *cur_y, const unsigned char *cur_u,
                                   const unsigned char *cur_v, unsigned char *rec_y, unsigned char *rec_u,
                                   unsigned char *rec_v, short *MB_out, int *modes_out, int *parts_out, int *segment_out,
                                   const int *quants) {
    memset(&video, 0, sizeof(video));
    video.wrk_width = width;
    video.wrk_height = height;
    video.mb_width = width / 16;
    video.mb_height = height / 16;
    video.mb_count = video.mb_width * video.mb_height;
    video.GOP_size = 1;  // (intra_transform()'s uploads are skipped: only the macroblock loop matters here)
    frames.current_Y = const_cast<unsigned char *>(cur_y);
    frames.current_U = const_cast<unsigned char *>(cur_u);
    frames.current_V = const_cast<unsigned char *>(cur_v);
    frames.reconstructed_Y = rec_y;
    frames.reconstructed_U = rec_u;
    frames.reconstructed_V = rec_v;
    frames.MB = (macroblock_coeffs_t *)calloc(video.mb_count, sizeof(macroblock_coeffs_t));
    frames.e_data = (macroblock_extra_data *)calloc(video.mb_count, sizeof(macroblock_extra_data));
    frames.MB_parts = parts_out;
    frames.MB_segment_id = segment_out;
    frames.y_dc_q[intra_segment] = quants[0];
    frames.y_ac_q[intra_segment] = quants[1];
    frames.uv_dc_q[intra_segment] = quants[2];
    frames.uv_ac_q[intra_segment] = quants[3];
    frames.frame_number = 0;
    intra_transform();
    memcpy(MB_out, frames.MB, (size_t)video.mb_count * sizeof(macroblock_coeffs_t));
    for (int mb = 0; mb < video.mb_count; ++mb)
        for (int b = 0; b < 16; ++b) modes_out[16 * mb + b] = frames.e_data[mb].mode[b];
    free(frames.MB);
    free(frames.e_data);
}
