"""ctypes mirrors of include/vp8_abi.h (binary-compatible with the reference's structs)."""
import ctypes as C

u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)


class KeyFrameHeader(C.Structure):
    """reference src/m02_vp8_header/vp8_header.h:7-18"""
    _fields_ = [
        ("is_key_frame", C.c_int), ("profile", C.c_uint8), ("show_frame", C.c_int),
        ("first_partition_len", C.c_uint32), ("start_code_ok", C.c_int),
        ("width", C.c_uint16), ("height", C.c_uint16), ("x_scale", C.c_uint8), ("y_scale", C.c_uint8),
    ]


class DecodedFrame(C.Structure):
    """reference src/m05_tokens/vp8_tokens.h:52-99"""
    _fields_ = [
        ("mb_cols", C.c_uint32), ("mb_rows", C.c_uint32), ("mb_total", C.c_uint32),
        ("q_index", C.c_uint8), ("y1_dc_delta_q", C.c_int8), ("y2_dc_delta_q", C.c_int8),
        ("y2_ac_delta_q", C.c_int8), ("uv_dc_delta_q", C.c_int8), ("uv_ac_delta_q", C.c_int8),
        ("segmentation_enabled", C.c_uint8), ("segmentation_abs", C.c_uint8),
        ("seg_quant_idx", C.c_int8 * 4), ("seg_lf_level", C.c_int8 * 4),
        ("lf_use_simple", C.c_uint8), ("lf_level", C.c_uint8), ("lf_sharpness", C.c_uint8),
        ("lf_delta_enabled", C.c_uint8), ("lf_ref_delta", C.c_int8 * 4), ("lf_mode_delta", C.c_int8 * 4),
        ("segment_id", u8p), ("skip_coeff", u8p), ("has_coeff", u8p), ("ymode", u8p), ("uv_mode", u8p),
        ("bmode", u8p),
        ("coeff_y2", i16p), ("coeff_y", i16p), ("coeff_u", i16p), ("coeff_v", i16p),
        ("stats_opaque", C.c_uint64 * 25),
    ]


class Yuv420Image(C.Structure):
    """reference src/m06_recon/vp8_recon.h:10-18"""
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("stride_y", C.c_uint32), ("stride_uv", C.c_uint32),
        ("y", u8p), ("u", u8p), ("v", u8p),
    ]


assert C.sizeof(KeyFrameHeader) == 28 and C.sizeof(DecodedFrame) == 320 and C.sizeof(Yuv420Image) == 40
