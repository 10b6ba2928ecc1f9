/* yolo3_b200_probe.h - hardware probes used by tests/probe_*.py to pin two sm_100a behaviours the convolution kernels
 * rely on (UMMA descriptors with row-shifted start addresses, im2col-mode TMA).  They are NOT part of the product:
 * they live in their own shared object, libyolo3_b200_probe.so, which links against libyolo3_b200.so and takes
 * handles created by it. */
#ifndef YOLO3_B200_PROBE_H_
#define YOLO3_B200_PROBE_H_

#include "yolo3_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Hardware probe (test hook): UMMA K-major SWIZZLE_128B descriptors with row-shifted start addresses.
 * a_bf16 [512][64] bf16 bits; out [2][n_shift][128][64] fp32: variant 0 = base_offset 0, variant 1 =
 * base_offset (addr >> 7) & 7; entry (v, i) should equal rows shifts[i] .. shifts[i]+127 of a. */
y3_status y3_debug_umma_rowshift(y3_handle h, const uint16_t* a_bf16, const int32_t* shifts, int32_t n_shift, float* out);

/* Hardware probe (test hook): im2col-mode TMA loads of 128 output pixels x 64 channels.  x_bf16 NHWC bf16 bits;
 * probes [n][6] = c, w, h, n, tap_w, tap_h; out [n][128][64] = the raw (128B-swizzled) shared-memory tiles. */
y3_status y3_debug_im2col(y3_handle h, const uint16_t* x_bf16, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride,
                          int32_t pad_lo, int32_t pad_hi, int32_t ksize, const int32_t* probes, int32_t n_probe, uint16_t* out);

#ifdef __cplusplus
}
#endif
#endif /* YOLO3_B200_PROBE_H_ */
