/*
 * vp8_abi.h - plain-C struct layouts shared across the drop-in boundary.
 *
 * These three structs are binary-compatible (x86-64 SysV) with the reference decoder's own
 * types, so that the reference's main.c / m01..m05 objects can be linked against libvp8gpu.so
 * unchanged:
 *
 *   Vp8KeyFrameHeader  <-> reference src/m02_vp8_header/vp8_header.h:7-18     (28 bytes)
 *   Vp8DecodedFrame    <-> reference src/m05_tokens/vp8_tokens.h:52-99        (320 bytes)
 *   Yuv420Image        <-> reference src/m06_recon/vp8_recon.h:10-18          (40 bytes)
 *
 * The 200-byte Vp8CoeffStats tail of Vp8DecodedFrame (vp8_tokens.h:7-50) is diagnostics only;
 * the pixel path never reads it, so it is carried as an opaque blob here.
 */
#ifndef VP8_ABI_H
#define VP8_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int is_key_frame;
	uint8_t profile;
	int show_frame;
	uint32_t first_partition_len;
	int start_code_ok;
	uint16_t width;  /* visible luma width, 1..16383  */
	uint16_t height; /* visible luma height, 1..16383 */
	uint8_t x_scale;
	uint8_t y_scale;
} Vp8KeyFrameHeader;

typedef struct {
	uint32_t mb_cols; /* ceil(width/16)  */
	uint32_t mb_rows; /* ceil(height/16) */
	uint32_t mb_total;

	/* quantiser indices (RFC 6386 9.6) */
	uint8_t q_index;
	int8_t y1_dc_delta_q;
	int8_t y2_dc_delta_q;
	int8_t y2_ac_delta_q;
	int8_t uv_dc_delta_q;
	int8_t uv_ac_delta_q;

	/* segmentation (RFC 6386 9.3) */
	uint8_t segmentation_enabled;
	uint8_t segmentation_abs;
	int8_t seg_quant_idx[4];
	int8_t seg_lf_level[4];

	/* loop filter (RFC 6386 9.6 / 15) */
	uint8_t lf_use_simple;
	uint8_t lf_level;
	uint8_t lf_sharpness;
	uint8_t lf_delta_enabled;
	int8_t lf_ref_delta[4];
	int8_t lf_mode_delta[4];

	/* per-macroblock syntax, raster order */
	uint8_t* segment_id; /* [mb_total] 0..3; read only when segmentation_enabled */
	uint8_t* skip_coeff; /* [mb_total]; not read by the pixel path               */
	uint8_t* has_coeff;  /* [mb_total] any non-zero coefficient; may be NULL     */
	uint8_t* ymode;      /* [mb_total] 0..4 = DC,V,H,TM,B_PRED                    */
	uint8_t* uv_mode;    /* [mb_total] 0..3 = DC,V,H,TM                           */
	uint8_t* bmode;      /* [mb_total*16]; read only where ymode==4               */

	/* dense coefficients, natural (de-zigzagged) order, 16 per 4x4 block */
	int16_t* coeff_y2; /* [mb_total*16]    */
	int16_t* coeff_y;  /* [mb_total*16*16] */
	int16_t* coeff_u;  /* [mb_total*4*16]  */
	int16_t* coeff_v;  /* [mb_total*4*16]  */

	uint64_t stats_opaque[25]; /* Vp8CoeffStats, untouched */
} Vp8DecodedFrame;

typedef struct {
	uint32_t width;
	uint32_t height;
	uint32_t stride_y;
	uint32_t stride_uv;
	uint8_t* y;
	uint8_t* u;
	uint8_t* v;
} Yuv420Image;

#if defined(__x86_64__) || defined(__aarch64__)
#ifdef __cplusplus
#define VP8_ABI_ASSERT(c, m) static_assert(c, m)
#else
#define VP8_ABI_ASSERT(c, m) _Static_assert(c, m)
#endif
VP8_ABI_ASSERT(sizeof(Vp8KeyFrameHeader) == 28, "Vp8KeyFrameHeader ABI");
VP8_ABI_ASSERT(offsetof(Vp8KeyFrameHeader, width) == 20, "Vp8KeyFrameHeader.width");
VP8_ABI_ASSERT(sizeof(Vp8DecodedFrame) == 320, "Vp8DecodedFrame ABI");
VP8_ABI_ASSERT(offsetof(Vp8DecodedFrame, segment_id) == 40, "Vp8DecodedFrame.segment_id");
VP8_ABI_ASSERT(offsetof(Vp8DecodedFrame, coeff_v) == 112, "Vp8DecodedFrame.coeff_v");
VP8_ABI_ASSERT(sizeof(Yuv420Image) == 40, "Yuv420Image ABI");
VP8_ABI_ASSERT(offsetof(Yuv420Image, y) == 16, "Yuv420Image.y");
#endif

#ifdef __cplusplus
}
#endif
#endif /* VP8_ABI_H */
