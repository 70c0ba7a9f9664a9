/*
 * vp8_parse.h - host-side front end feeding the GPU pixel path: RIFF/WebP container, VP8 key-frame header,
 * per-macroblock modes and DCT tokens -> Vp8DecodedFrame (include/vp8_abi.h).
 *
 * Stands in for the reference's m01_container + m02_vp8_header + m05_tokens
 * (webp_container.c:19-89, vp8_header.c:13-65, vp8_tokens.c:673-1001) with two differences that the GPU
 * pipeline needs: it is RE-ENTRANT (the reference keeps its coefficient probabilities in a writable global,
 * vp8_tokens.c:625) so frames can be parsed one per host thread, and it can write its arrays into caller
 * memory (e.g. pinned staging) instead of calloc. Entropy decoding is inherently serial and stays on the CPU.
 *
 * Same acceptance rules and errors as the reference: simple lossy files only (RIFF/WEBP + one 'VP8 ' chunk,
 * sizes consistent), key frames only, a single token partition (ENOTSUP otherwise); -1 with errno set.
 */
#ifndef VP8_PARSE_H
#define VP8_PARSE_H

#include "vp8_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Bytes of array storage a frame of this size needs (all ten arrays, each 64-byte aligned). */
size_t vp8_parse_arena_bytes(uint32_t width, uint32_t height);

/* Reads just enough of a .webp file to report the frame size. */
int vp8_parse_webp_size(const uint8_t* file, size_t size, uint32_t* width, uint32_t* height);

/* Parse one .webp file. With arena == NULL the arrays are calloc'ed (release with vp8_parse_free); otherwise
 * they are carved from arena[0..arena_bytes) (>= vp8_parse_arena_bytes), which the caller owns. */
int vp8_parse_webp(const uint8_t* file, size_t size, Vp8KeyFrameHeader* kf, Vp8DecodedFrame* out, void* arena,
                   size_t arena_bytes);

/* Same for a bare 'VP8 ' chunk payload. */
int vp8_parse_vp8(const uint8_t* payload, size_t size, Vp8KeyFrameHeader* kf, Vp8DecodedFrame* out, void* arena,
                  size_t arena_bytes);

void vp8_parse_free(Vp8DecodedFrame* f);

/* n files on `threads` host threads (one image per thread at a time). arenas[i] may be NULL (calloc).
 * status[i] receives 0 or the errno of file i; returns the number of failed files. */
int vp8_parse_batch(const uint8_t* const* files, const size_t* sizes, int n, int threads, Vp8KeyFrameHeader* kf,
                    Vp8DecodedFrame* out, void* const* arenas, const size_t* arena_bytes, int* status);

#ifdef __cplusplus
}
#endif
#endif /* VP8_PARSE_H */
