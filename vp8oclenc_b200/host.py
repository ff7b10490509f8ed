"""Python host-side mirror of the kernel-level C ABI (include/vp8b200.h).

One function per reference __kernel, same names and argument meaning as the reference's
kernels (src/GPU_kernels.cl, src/CPU_kernels.cl).  Arguments are torch CUDA tensors (torch is
only used for device memory and streams); every call is forwarded to libvp8b200.so, which
contains nothing but CUDA code -- there is no CPU fallback: loading fails loudly when the
library has not been built, and every call raises on a CUDA error.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("VP8B200_ENGINE_LIB") or os.path.join(_PKG, "lib", "libvp8b200.so")  # (the variable: kernel-tuning variants, tools/)
_lib = None


class EngineError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise EngineError("%s is missing: run `python -m vp8oclenc_b200.build` (there is no fallback path)" % _LIB_PATH)
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.vp8b200_version.restype = ctypes.c_char_p
    return _lib


def _p(t):
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        raise EngineError("vp8oclenc_b200 kernels take CUDA tensors only")
    if not t.is_contiguous():
        raise EngineError("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, what):
    if rc != 0:
        raise EngineError("%s failed: CUDA error %d" % (what, -rc))


def device_info():
    name = ctypes.create_string_buffer(256)
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(lib().vp8b200_device_info(name, 256, ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "device_info")
    return dict(name=name.value.decode(), sm_count=sm.value, cc=(major.value, minor.value))


def segment_data_tensor(sd_np):
    """numpy int32 [4, 11] -> device copy"""
    return torch.from_numpy(sd_np.copy()).cuda()


def reset_vectors(last1, last2, gold1, gold2, alt1, alt2, last_d, gold_d, alt_d):
    _check(lib().vp8b200_reset_vectors(_stream(), _p(last1), _p(last2), _p(gold1), _p(gold2), _p(alt1), _p(alt2),
                                       _p(last_d), _p(gold_d), _p(alt_d), ctypes.c_int(last_d.numel())), "reset_vectors")


def downsample_x2(src, dst, src_width, src_height):
    _check(lib().vp8b200_downsample_x2(_stream(), _p(src), _p(dst), src_width, src_height), "downsample_x2")


def luma_search_1step(current_frame, prev_frame, src_net, dst_net, net_width, width, height, pixel_rate):
    _check(lib().vp8b200_luma_search_1step(_stream(), _p(current_frame), _p(prev_frame), _p(src_net), _p(dst_net),
                                           net_width, width, height, pixel_rate), "luma_search_1step")


def luma_search_2step(current_frame, ref_frame, net, ref_net, ref_Bdiff, width, height):
    _check(lib().vp8b200_luma_search_2step(_stream(), _p(current_frame), _p(ref_frame), _p(net), _p(ref_net),
                                           _p(ref_Bdiff), width, height), "luma_search_2step")


def select_reference(last_net, golden_net, altref_net, last_d, golden_d, altref_d, MB_reference_frame, MB_vectors,
                     width, height, use_golden, use_altref):
    _check(lib().vp8b200_select_reference(_stream(), _p(last_net), _p(golden_net), _p(altref_net), _p(last_d),
                                          _p(golden_d), _p(altref_d), _p(MB_reference_frame), _p(MB_vectors), width,
                                          height, int(use_golden), int(use_altref)), "select_reference")


def pack_8x8_into_16x16(MB_vectors, MB_parts, MB_SSIM):
    _check(lib().vp8b200_pack_8x8_into_16x16(_stream(), _p(MB_vectors), _p(MB_parts), _p(MB_SSIM),
                                             ctypes.c_int(MB_parts.numel())), "pack_8x8_into_16x16")


def prepare_predictors_and_residual(current_frame, ref_frame, predictor, residual, MB_reference_frame, MB_vectors,
                                    width, height, plane, ref):
    _check(lib().vp8b200_prepare_predictors_and_residual(_stream(), _p(current_frame), _p(ref_frame), _p(predictor),
                                                         _p(residual), _p(MB_reference_frame), _p(MB_vectors), width,
                                                         height, plane, ref), "prepare_predictors_and_residual")


def dct4x4(residual, MB, MB_segment_id, MB_parts, MB_SSIM, width, height, SD, segment_id, SSIM_target, plane):
    _check(lib().vp8b200_dct4x4(_stream(), _p(residual), _p(MB), _p(MB_segment_id), _p(MB_parts), _p(MB_SSIM), width,
                                height, _p(SD), segment_id, ctypes.c_float(SSIM_target), plane), "dct4x4")


def wht4x4_iwht4x4(MB, MB_segment_id, MB_parts, SD, segment_id):
    _check(lib().vp8b200_wht4x4_iwht4x4(_stream(), _p(MB), _p(MB_segment_id), _p(MB_parts), _p(SD), segment_id,
                                        ctypes.c_int(MB_parts.numel())), "wht4x4_iwht4x4")


def idct4x4(recon_frame, predictor, MB, MB_segment_id, MB_parts, width, height, SD, segment_id, plane):
    _check(lib().vp8b200_idct4x4(_stream(), _p(recon_frame), _p(predictor), _p(MB), _p(MB_segment_id), _p(MB_parts),
                                 width, height, _p(SD), segment_id, plane), "idct4x4")


def count_SSIM(frame1, frame2, MB_segment_id, metric, width, height, segment_id, mb_size):
    _check(lib().vp8b200_count_SSIM(_stream(), _p(frame1), _p(frame2), _p(MB_segment_id), _p(metric), width, height,
                                    segment_id, mb_size), "count_SSIM")


def gather_SSIM(metric1, metric2, metric3, MB_SSIM):
    _check(lib().vp8b200_gather_SSIM(_stream(), _p(metric1), _p(metric2), _p(metric3), _p(MB_SSIM),
                                     ctypes.c_int(MB_SSIM.numel())), "gather_SSIM")


def prepare_filter_mask(MB, MB_non_zero_coeffs, MB_parts, mb_mask, width, height):
    _check(lib().vp8b200_prepare_filter_mask(_stream(), _p(MB), _p(MB_non_zero_coeffs), _p(MB_parts), _p(mb_mask),
                                             width, height), "prepare_filter_mask")


def loop_filter_frame(frame, MB_segment_ids, mb_mask, SD, width, height, mb_size):
    _check(lib().vp8b200_loop_filter_frame(_stream(), _p(frame), _p(MB_segment_ids), _p(mb_mask), _p(SD), width,
                                           height, mb_size), "loop_filter_frame")


def loop_filter_planes(y, u, v, MB_segment_ids, mb_mask, SD, width, height):
    _check(lib().vp8b200_loop_filter_planes(_stream(), _p(y), _p(u), _p(v), _p(MB_segment_ids), _p(mb_mask), _p(SD),
                                            width, height), "loop_filter_planes")


def entropy_tokens(MB, MB_non_zero_coeffs, MB_parts, mb_width, mb_height, num_partitions, coeff_probs, coeff_probs_denom,
                   third_context, tokens, mb_tokens, mb_offset, part_info, tail_scratch):
    """GPU half of count_probs + encode_coefficients (src/CPU_kernels.cl:347-778): statistics, contexts and the
    decision stream of every partition; `tokens` is an int16/uint16 tensor whose length is the capacity"""
    _check(lib().vp8b200_entropy_tokens(_stream(), _p(MB), _p(MB_non_zero_coeffs), _p(MB_parts), mb_width, mb_height,
                                        num_partitions, _p(coeff_probs), _p(coeff_probs_denom), _p(third_context),
                                        _p(tokens), ctypes.c_uint32(tokens.numel()), _p(mb_tokens), _p(mb_offset),
                                        _p(part_info), _p(tail_scratch)), "entropy_tokens")


def entropy_boolcode(tokens, part_info, coeff_probs, output, partition_sizes, num_partitions, partition_step, max_decisions):
    """encode_coefficients (src/CPU_kernels.cl:541-778) over the decision streams of entropy_tokens: RFC 6386's
    boolean coder on the GPU in parallel; partition p lands at output[p * partition_step:].  max_decisions: upper
    bound of the decisions of one partition (the total of part_info will do)"""
    import torch
    L = lib()
    L.vp8b200_entropy_boolcode_scratch_bytes.restype = ctypes.c_size_t
    need = L.vp8b200_entropy_boolcode_scratch_bytes(ctypes.c_uint32(max_decisions), num_partitions, partition_step)
    scratch = torch.empty(need, dtype=torch.uint8, device=tokens.device)
    _check(L.vp8b200_entropy_boolcode(_stream(), _p(tokens), _p(part_info), _p(coeff_probs), _p(output),
                                      _p(partition_sizes), num_partitions, partition_step,
                                      ctypes.c_uint32(max_decisions), _p(scratch)), "entropy_boolcode")
    return scratch  # (keep it alive until the stream has run)


def intra_frame(cur_y, cur_u, cur_v, rec_y, rec_u, rec_v, MB, modes, MB_parts, MB_segment_id, width, height, quants):
    """intra_transform() / predict_and_transform_mb() of the reference host for a whole frame (src/intra_part.h:37-741,
    1089-1128): B_PRED luma, TM_PRED chroma, transform, quantise, reconstruct.  quants = (y_dc_q, y_ac_q, uv_dc_q, uv_ac_q)"""
    import torch
    L = lib()
    L.vp8b200_intra_frame_scratch_bytes.restype = ctypes.c_size_t
    scratch = torch.empty(L.vp8b200_intra_frame_scratch_bytes(width, height), dtype=torch.uint8, device=cur_y.device)
    _check(L.vp8b200_intra_frame(_stream(), _p(cur_y), _p(cur_u), _p(cur_v), _p(rec_y), _p(rec_u), _p(rec_v), _p(MB), _p(modes),
                                 _p(MB_parts), _p(MB_segment_id), width, height, int(quants[0]), int(quants[1]), int(quants[2]),
                                 int(quants[3]), _p(scratch)), "intra_frame")
    return scratch  # (keep it alive until the stream has run)


# ---------------------------------------------------------------------------------------------
class Engine:
    """Frame-level engine (vp8b200_engine_* of include/vp8b200.h): the sequence of
    prepare_GPU_buffers() + inter_transform() + loop filter of the reference host
    (src/inter_part.h, src/loop_filter.h) on one CUDA stream."""

    BUF = dict(coeffs=0, vectors=1, parts=2, reference_frame=3, segment_id=4, ssim=5, non_zero=6, recon_y=7,
               recon_u=8, recon_v=9, intra_modes=10)

    def __init__(self, width, height, stream=None):
        L = lib()
        L.vp8b200_engine_create.restype = ctypes.c_void_p
        L.vp8b200_engine_buffer.restype = ctypes.c_void_p
        L.vp8b200_engine_stream.restype = ctypes.c_void_p
        self.width, self.height = width, height
        self.mb_count = (width // 16) * (height // 16)
        h = L.vp8b200_engine_create(width, height, ctypes.c_void_p(stream or 0))
        if not h:
            raise EngineError("vp8b200_engine_create failed (no CUDA device, or size not a multiple of 16)")
        self._h = ctypes.c_void_p(h)

    def close(self):
        if self._h:
            lib().vp8b200_engine_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def stream(self):
        return lib().vp8b200_engine_stream(self._h)

    def synchronize(self):
        _check(lib().vp8b200_engine_synchronize(self._h), "engine_synchronize")

    def set_reconstruction(self, y, u, v):
        on_dev = int(y.is_cuda)
        ptr = (lambda t: _p(t)) if on_dev else (lambda t: ctypes.c_void_p(t.data_ptr()))
        _check(lib().vp8b200_engine_set_reconstruction(self._h, ptr(y), ptr(u), ptr(v), on_dev), "set_reconstruction")

    def inter_frame(self, cur_y, cur_u, cur_v, sd_np, ssim_target, prev_is_golden, prev_is_altref, altref_differs):
        _check(lib().vp8b200_engine_inter_frame(self._h, _p(cur_y), _p(cur_u), _p(cur_v),
                                                ctypes.c_void_p(sd_np.ctypes.data), ctypes.c_float(ssim_target),
                                                int(prev_is_golden), int(prev_is_altref), int(altref_differs)),
               "engine_inter_frame")

    def key_frame(self, cur_y, cur_u, cur_v, sd_np):
        """intra_transform() for the current frame (device tensors); follow with loop_filter(sd_np)"""
        _check(lib().vp8b200_engine_key_frame(self._h, _p(cur_y), _p(cur_u), _p(cur_v), ctypes.c_void_p(sd_np.ctypes.data)),
               "engine_key_frame")

    def loop_filter(self, sd_np=None):
        p = ctypes.c_void_p(sd_np.ctypes.data) if sd_np is not None else ctypes.c_void_p(0)
        _check(lib().vp8b200_engine_loop_filter(self._h, p), "engine_loop_filter")

    def encode_frame_host(self, cur_y, cur_u, cur_v, sd_np, ssim_target, prev_is_golden, prev_is_altref, altref_differs,
                          out):
        """cur_* and the tensors in `out` (dict keyed like BUF) are HOST tensors (pinned for speed)"""
        def hp(name):
            t = out.get(name)
            return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
        _check(lib().vp8b200_engine_encode_frame_host(
            self._h, ctypes.c_void_p(cur_y.data_ptr()), ctypes.c_void_p(cur_u.data_ptr()),
            ctypes.c_void_p(cur_v.data_ptr()), ctypes.c_void_p(sd_np.ctypes.data), ctypes.c_float(ssim_target),
            int(prev_is_golden), int(prev_is_altref), int(altref_differs), hp("coeffs"), hp("vectors"), hp("parts"),
            hp("reference_frame"), hp("segment_id"), hp("ssim"), hp("non_zero"), hp("recon_y"), hp("recon_u"),
            hp("recon_v")), "engine_encode_frame_host")

    def buffer_ptr(self, name):
        return lib().vp8b200_engine_buffer(self._h, self.BUF[name])

    def read(self, name):
        """device result buffer -> numpy (synchronising)"""
        import numpy as np
        M, w, h = self.mb_count, self.width, self.height
        shapes = dict(coeffs=(np.int16, M * 400), vectors=(np.int16, M * 8), parts=(np.int32, M),
                      reference_frame=(np.int32, M), segment_id=(np.int32, M), ssim=(np.float32, M),
                      non_zero=(np.int32, M), recon_y=(np.uint8, w * h), recon_u=(np.uint8, w * h // 4),
                      recon_v=(np.uint8, w * h // 4), intra_modes=(np.int32, M * 16))
        dt, n = shapes[name]
        out = np.empty(n, dt)
        self.synchronize()
        rc = _cudart().cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(self.buffer_ptr(name)),
                                  ctypes.c_size_t(out.nbytes), 2)
        if rc != 0:
            raise EngineError("cudaMemcpy D2H failed: %d" % rc)
        return out

    @property
    def last_launch_count(self):
        return lib().vp8b200_engine_last_launch_count(self._h)

    STAGES = ("setup", "search_16x", "search_8x", "search_4x", "search_2x", "search_1x", "search_qpel", "select",
              "transform", "filter_mask", "loop_filter")

    def stage_timing(self, on=True):
        _check(lib().vp8b200_engine_stage_timing(self._h, int(on)), "engine_stage_timing")

    def stage_times(self):
        """{stage name: milliseconds} of the last frame (waits for the stream); stages that did not run are left out"""
        buf = (ctypes.c_float * len(self.STAGES))()
        _check(lib().vp8b200_engine_stage_times(self._h, buf, len(self.STAGES)), "engine_stage_times")
        return {n: float(buf[i]) for i, n in enumerate(self.STAGES) if buf[i] >= 0.0}


_cudart_lib = None


def _cudart():
    global _cudart_lib
    if _cudart_lib is None:
        import glob
        cands = glob.glob("/usr/local/cuda/lib64/libcudart.so*") + glob.glob("/usr/local/cuda/targets/*/lib/libcudart.so*")
        _cudart_lib = ctypes.CDLL(sorted(cands)[0]) if cands else ctypes.CDLL("libcudart.so")
    return _cudart_lib
