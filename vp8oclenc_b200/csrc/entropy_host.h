// Host-side execution of the three sequential "CPU program" kernels that produce the
// coefficient partitions of the VP8 frame:
//   count_probs, num_div_denom, encode_coefficients   (src/CPU_kernels.cl:347-410, 541-778)
// They are boolean-coder / token-statistics code with a serial dependency inside each
// partition, so they run on host threads, one per partition, exactly as the reference runs
// them on its CPU OpenCL device (src/vp8enc.cpp:65-88).  Outputs are bit-identical to the
// reference's: token statistics (including its habit of counting end-of-block decisions for
// every position after the first end-of-block), probabilities, contexts and partition bytes.
#pragma once
#include <stdint.h>

namespace vp8host {

// coeff_probs / coeff_probs_denom: uint32 [partitions][4][8][3][11]; third_context: uint8 [mb][25]
void count_probs(const int16_t *MB, const int32_t *MB_non_zero_coeffs, const int32_t *MB_parts, uint32_t *coeff_probs,
                 uint32_t *coeff_probs_denom, uint8_t *third_context, int mb_height, int mb_width, int num_partitions);

void num_div_denom(uint32_t *coeff_probs, const uint32_t *coeff_probs_denom, int num_partitions);

void encode_coefficients(const int16_t *MB, const int32_t *MB_non_zero_coeffs, const int32_t *MB_parts,
                         uint8_t *output, int32_t *partition_sizes, const uint8_t *third_context,
                         const uint32_t *coeff_probs, int mb_height, int mb_width, int num_partitions,
                         int partition_step);

// bool-codes the decision streams prepared on the GPU (vp8b200_entropy_tokens); same bytes as
// encode_coefficients
void encode_token_streams(const uint16_t *tokens, const uint32_t *part_info, const uint32_t *coeff_probs, uint8_t *output,
                          int32_t *partition_sizes, int num_partitions, int partition_step);

}  // namespace vp8host
