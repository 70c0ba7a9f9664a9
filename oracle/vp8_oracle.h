/*
 * vp8_oracle.h - CPU restatement of the reference pixel path (m06 recon, m07 loop filter,
 * m08 YUV->RGB, m09 PNG framing).
 *
 * TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg are
 * the only callers. The product library (webp-decoder_b200/csrc) neither links nor loads it.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against the real
 * reference (oracle/_ref/libref_hot.so, built from /root/reference by oracle/Makefile) on the
 * fixture corpus and on struct-level fuzz frames, and tests/golden/ holds digests produced by
 * that reference which are re-checked wherever /root/reference is absent.
 */
#ifndef VP8_ORACLE_H
#define VP8_ORACLE_H

#include "../include/vp8_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Padded (mb_cols*16 x mb_rows*16, chroma half) planes; strides = padded widths. */
int orc_recon_padded(const Vp8DecodedFrame* f, uint8_t* y, uint8_t* u, uint8_t* v);
int orc_loopfilter_padded(const Vp8DecodedFrame* f, uint8_t* y, uint8_t* u, uint8_t* v);

/* {y1dc,y1ac,uvdc,uvac,y2dc,y2ac} per segment and {level,interior,hev,0} per segment x (ymode==B_PRED). */
void orc_frame_params(const Vp8DecodedFrame* f, int16_t dq[4][6], uint8_t lf[4][2][4]);

/* Full path to a tight I420 buffer: Y[w*h] U[cw*ch] V[cw*ch], cw=(w+1)/2, ch=(h+1)/2. */
int orc_decode_i420(const Vp8DecodedFrame* f, uint32_t width, uint32_t height, int filtered, uint8_t* out);

/* Fancy-upsampled RGB24, tight (w*3 bytes per row). */
void orc_i420_to_rgb(const uint8_t* y, const uint8_t* u, const uint8_t* v, uint32_t width, uint32_t height,
                     uint32_t stride_y, uint32_t stride_uv, uint8_t* rgb);

/* "P6\n<w> <h>\n255\n" + RGB. Returns bytes written (out must hold 32 + w*h*3). */
size_t orc_ppm(const uint8_t* rgb, uint32_t width, uint32_t height, uint8_t* out);

/* Stored-deflate PNG exactly as the reference frames it. orc_png_bound gives the buffer size. */
size_t orc_png_bound(uint32_t width, uint32_t height);
size_t orc_png(const uint8_t* rgb, uint32_t width, uint32_t height, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
