"""int32 op-list records consumed by ``sb_load_model`` (layout mirrored in csrc/sb_model.h)."""
import struct

import numpy as np

SB_OP_WORDS = 24
BUFFER, CONV, TCONV, POOL, UPSAMPLE, ADD, PREPROCESS, COPY = 0, 1, 2, 3, 4, 5, 6, 7
F_RELU, F_BN, F_BILINEAR, F_FUSED_POOL = 1, 2, 8, 16


def _rec():
    r = np.zeros((SB_OP_WORDS,), np.int32)
    r[4] = -1
    r[12:16] = -1
    r[18] = -1
    return r


def buffer_record(buf_id, stride_den, C, f32, is_input):
    r = np.zeros((SB_OP_WORDS,), np.int32)
    r[0], r[1], r[2], r[3], r[4], r[5] = BUFFER, buf_id, stride_den, C, int(bool(f32)), int(is_input)
    return r


def preprocess_record(out_buf, C, input_scale, pad_stride):
    r = _rec()
    r[0], r[1], r[6], r[8] = PREPROCESS, -1, out_buf, C
    r[16] = struct.unpack("<i", struct.pack("<f", float(input_scale)))[0]
    r[17] = int(pad_stride)
    return r


def conv_record(in_buf, in_coff, in_C, out_buf, out_coff, out_C, k, stride, relu, w_off, b_off,
                bn_scale_off=-1, bn_shift_off=-1, pool_buf=-1, pool_coff=0):
    r = _rec()
    r[0], r[1], r[2], r[3] = CONV, in_buf, in_coff, in_C
    r[6], r[7], r[8], r[9], r[10] = out_buf, out_coff, out_C, k, stride
    r[11] = (F_RELU if relu else 0) | (F_BN if bn_scale_off >= 0 else 0)
    r[12], r[13], r[14], r[15] = w_off, b_off, bn_scale_off, bn_shift_off
    r[18], r[19] = pool_buf, pool_coff
    return r


def tconv_record(in_buf, in_coff, in_C, out_buf, out_coff, out_C, w_off, b_off):
    r = _rec()
    r[0], r[1], r[2], r[3] = TCONV, in_buf, in_coff, in_C
    r[6], r[7], r[8], r[9], r[10] = out_buf, out_coff, out_C, 3, 2
    r[11] = F_RELU
    r[12], r[13] = w_off, b_off
    return r


def pool_record(in_buf, in_coff, C, out_buf, out_coff, fused=False):
    r = _rec()
    r[0], r[1], r[2], r[3], r[6], r[7], r[8] = POOL, in_buf, in_coff, C, out_buf, out_coff, C
    r[11] = F_FUSED_POOL if fused else 0
    return r


def upsample_record(in_buf, in_coff, C, out_buf, out_coff, bilinear):
    r = _rec()
    r[0], r[1], r[2], r[3], r[6], r[7], r[8] = UPSAMPLE, in_buf, in_coff, C, out_buf, out_coff, C
    r[11] = F_BILINEAR if bilinear else 0
    return r


def add_record(a_buf, a_coff, b_buf, b_coff, C, out_buf, out_coff):
    r = _rec()
    r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8] = ADD, a_buf, a_coff, C, b_buf, b_coff, out_buf, out_coff, C
    return r


def copy_record(in_buf, in_coff, C, out_buf, out_coff):
    r = _rec()
    r[0], r[1], r[2], r[3], r[6], r[7], r[8] = COPY, in_buf, in_coff, C, out_buf, out_coff, C
    return r
