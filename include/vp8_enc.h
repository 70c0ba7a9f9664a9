/*
 * vp8_enc.h - C-ABI of the encoder-side neighbour of the pixel path: the reference ENCODER's in-loop
 * reconstruction, on the GPU. Per macroblock, in dependency order: predict from the reconstructed neighbours (DC; the best
 * of DC / V / H / TM by squared error; or per 4x4 sub-block the best of the ten B_PRED modes by SAD), forward DCT (+ WHT),
 * quantise, then the decoder's dequantise / inverse WHT / inverse DCT / add - the same device transforms and predictors as
 * the decoder kernels, run as a macroblock wavefront, many pictures per launch.
 *
 * Two groups, as in vp8_gpu.h:
 *  (1) the reference's own entry points, same prototypes, ownership (callee-malloc'ed arrays, caller frees) and error
 *      convention (0 / -1 + errno); linking the reference's unmodified encoder_main.c + modules against libvp8gpu.so in
 *      place of these functions gives byte-identical .webp files for --mode dc, --mode i16 and --mode bpred;
 *  (2) a batch entry point with caller-owned output buffers.
 * No CPU fallback: without a CUDA device every call fails with EIO.
 *
 * Built: the DC, the whole-macroblock (--mode i16) and the 4x4 sub-block SAD (--mode bpred) front ends. Not built: the RDO
 * flavour of the sub-block search (--mode bpred-rdo, enc_recon.c:1087-1187, 1833-2607), whose score needs the token-cost model
 * of enc-m07_tokens.
 */
#ifndef VP8_ENC_H
#define VP8_ENC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* binary compatible with reference src/enc-m04_yuv/enc_rgb_to_yuv.h:9-17 */
typedef struct EncYuv420Image {
	uint32_t width;
	uint32_t height;
	uint32_t y_stride;
	uint32_t uv_stride;
	uint8_t* y;
	uint8_t* u;
	uint8_t* v;
} EncYuv420Image;

/* ---------------------------------------------------------------- (1) reference module interface */

/* replaces reference src/enc-m08_recon/enc_recon.h:69-73 (enc_recon.c:855-1085): DC prediction for luma and chroma.
 * coeffs_out: mb_total * 400 int16, per macroblock Y2[16] Y[16][16] U[4][16] V[4][16], natural order. */
int enc_vp8_encode_dc_pred_inloop(const EncYuv420Image* yuv, int quality, int16_t** coeffs_out, size_t* coeffs_count_out,
                                  uint8_t* qindex_out);

/* replaces reference src/enc-m08_recon/enc_recon.h:97-105 (enc_recon.c:1189-1483): luma and chroma mode each the best of
 * DC=0, V=1, H=2, TM=3 by squared error against predictors built from reconstructed neighbours. */
int enc_vp8_encode_i16x16_uv_sad_inloop(const EncYuv420Image* yuv, int quality, uint8_t** y_modes_out, size_t* y_modes_count_out,
                                        uint8_t** uv_modes_out, size_t* uv_modes_count_out, int16_t** coeffs_out,
                                        size_t* coeffs_count_out, uint8_t* qindex_out);

/* replaces reference src/enc-m08_recon/enc_recon.h:83-89 (enc_recon.c:1485-1505): the same, chroma modes not returned */
int enc_vp8_encode_i16x16_sad_inloop(const EncYuv420Image* yuv, int quality, uint8_t** y_modes_out, size_t* y_modes_count_out,
                                     int16_t** coeffs_out, size_t* coeffs_count_out, uint8_t* qindex_out);

/* replaces reference src/enc-m08_recon/enc_recon.h:120-130 (enc_recon.c:1507-1831): every macroblock B_PRED (y_modes all 4);
 * per 4x4 sub-block the best of the ten modes by SAD against predictors built from the reconstruction so far (b_modes:
 * mb_total * 16, values 0..9), chroma mode by SAD; no Y2 block (its sixteen coefficients are zero). */
int enc_vp8_encode_bpred_uv_sad_inloop(const EncYuv420Image* yuv, int quality, uint8_t** y_modes_out, size_t* y_modes_count_out,
                                       uint8_t** b_modes_out, size_t* b_modes_count_out, uint8_t** uv_modes_out, size_t* uv_modes_count_out,
                                       int16_t** coeffs_out, size_t* coeffs_count_out, uint8_t* qindex_out);

/* ---------------------------------------------------------------- (2) batch interface */

/* n pictures in one launch on `device` (-1: VP8_GPU_DEVICE or 0). search = 0: DC prediction (dc_pred_inloop), 1: mode
 * search (i16x16_uv_sad_inloop). Per picture i, caller-owned: coeffs[i] (vp8_gpu_enc_mb_total(w, h) * 400 int16),
 * y_modes[i] / uv_modes[i] (mb_total bytes each; the arrays or single entries may be NULL), rec_y/u/v[i] (the macroblock-
 * aligned reconstruction, stride 16 * mb_cols / 8 * mb_cols; arrays or entries may be NULL). Pinned host memory
 * (vp8_gpu_host_alloc) makes the copies asynchronous. qindex_out: one byte, the same for every picture. */
int vp8_gpu_enc_i16_inloop(int device, const EncYuv420Image* const* yuv, int n, int quality, int search, int16_t* const* coeffs,
                           uint8_t* const* y_modes, uint8_t* const* uv_modes, uint8_t* const* rec_y, uint8_t* const* rec_u,
                           uint8_t* const* rec_v, uint8_t* qindex_out);
/* The sub-block front end for n pictures: as above plus b_modes[i] (mb_total * 16 bytes each, required). */
int vp8_gpu_enc_bpred_inloop(int device, const EncYuv420Image* const* yuv, int n, int quality, int16_t* const* coeffs,
                             uint8_t* const* y_modes, uint8_t* const* b_modes, uint8_t* const* uv_modes, uint8_t* const* rec_y,
                             uint8_t* const* rec_u, uint8_t* const* rec_v, uint8_t* qindex_out);
size_t vp8_gpu_enc_mb_total(uint32_t width, uint32_t height);
/* device time of the kernel of the last vp8_gpu_enc_*_inloop call of this thread, in milliseconds */
double vp8_gpu_enc_last_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif
