// On-disk / in-memory layout of a packed stage network ("blob").  Written by
// cnn_av1_research_b200/packer.py, read by av1p_model_create.  Little-endian, all data offsets
// are byte offsets from the start of the blob and 256-byte aligned.
//
// A blob is a tiny program over fp16 activation buffers: the packer decides the topology
// (which reference layer reads / writes which buffer); the runtime only interprets ops.
#pragma once
#include <stdint.h>

#define AV1P_BLOB_MAGIC 0x50315641u   /* "AV1P" */
#define AV1P_BLOB_VERSION 12u
#define AV1P_BLOB_MAX_NT 8
#define AV1P_BLOB_MAX_KB 128

enum Av1pOpType : int32_t {
  AV1P_OP_STEM = 0,       // gather + /1023 + conv1/bn/relu/maxpool -> out buffer (1024 cols); w = fp16 [4][128][64]: planes 0,1 =
                          // hi/lo weights for float blocks (f0 = acc_scale), planes 2,3 = hi/lo of w / 1023 for integer
                          // frame samples (f1 = acc_scale), planes 4,5 = hi/lo of w / 1023 * 2^s for raw uint16 words read
                          // as fp16 (sample * 2^-24), K = ky * 8 + kx + 1 (stem_tma.cuh; tail_n = 24 - s: acc_scale = 2^tail_n)
  AV1P_OP_FC = 1,         // block-Toeplitz linear layer on tensor cores
  AV1P_OP_SAM = 2,        // spatial-attention scalar of `src0` (512 cols) -> row_scale; tail_n = 1: from the partials of the
                          // preceding FC op (use_row_scale bit 2) instead of a pass over src0
  AV1P_OP_FGVC_TAIL = 3,  // L2-normalise `src0` (512 cols) + cosine classifier -> logits[4]
  AV1P_OP_SE = 4,         // squeeze-excite on CUDA cores: block_n = channels (64/128/256), n_tiles = positions
  AV1P_OP_CONV_RES = 5,   // 3x3 s1 conv 64->64 on the 4x4 map with SMEM-resident weights (csrc/conv_res_tcgen05.cuh):
                          // src = {x_hi, x_lo}, w = fp16 [planes][ky][kx = 2,1,0][64 co][64 ci], bias[1024], f0 = acc_scale,
                          // pair_mode = 1 for hi/lo planes (three products)
  // generic block sizes (8 / 32 / 64; csrc/gen_kernels.cuh) - every conv / linear layer stays an AV1P_OP_FC
  AV1P_OP_STEM_GEN = 6,   // gather + /1023 + conv1/bn/relu/maxpool in fp32: n_tiles = block size, w = fp32 [64][49], bias[64]
  AV1P_OP_SE_GEN = 7,     // squeeze-excite, any shape: block_n = channels, n_tiles = positions, w as AV1P_OP_SE
  AV1P_OP_SAM_POOL = 8,   // spatial attention (7x7 conv over [mean_c, max_c]) + global average pool: src0 = [g*g][512],
                          // n_tiles = g, w = fp32 [2][49] -> out (512 cols)
};

#pragma pack(push, 1)
struct Av1pBlobHeader {
  uint32_t magic, version;
  uint32_t stage_kind;    // informational (0 stage1, 1 stage2, 2 rect, 3 ab-fgvc, 4 ab-plain, 5 flat7)
  uint32_t n_ops, n_bufs, n_out;
  uint32_t precision;     // 0 fp16x3, 1 fp16 (informational)
  uint32_t block_size;    // luma block size the program was packed for: 8 / 16 / 32 / 64 (0 reads as 16)
  uint64_t ops_off, bufs_off, total_bytes;
  uint64_t reserved2;
};  // 64 bytes

struct Av1pBlobOp {
  int32_t type;
  int32_t src[4];                   // activation sources (buffer ids, -1 = none).  Split precision: {x_hi, x_lo, y_hi, y_lo}
  int32_t aux, aux_lo, out, out_lo; // buffer ids, -1 = none
  int32_t n_tiles, block_n, epi, tail_n;
  int32_t use_row_scale;            // bit 0: scale the accumulator with the spatial-attention scalar; bit 1: scale the aux operand
                                    // (FC_EPI_ADD); bit 2: leave the spatial-attention partials of the output behind (se4.fc2)
  int32_t n_kb_total;               // schedule entries
  int32_t n_w_chunks;               // [block_n x 64] fp16 weight tiles stored at w_off
  int32_t pair_mode;                // 1: entries are (x_hi,w_hi),(x_lo,w_lo) pairs; three products per pair
  float f0, f1;                     // SAM: w_avg, w_max.  FGVC tail: scale.  FC: f0 = acc_scale
  uint64_t w_off, bias_off, tail_w_off, tail_b_off;   // 0 = absent
  int32_t kb_begin[AV1P_BLOB_MAX_NT + 1];
  uint16_t kb_src[AV1P_BLOB_MAX_KB];   // bits 14..15: index into src[], bits 0..13: K offset / 64
  uint16_t kb_w[AV1P_BLOB_MAX_KB];     // weight chunk index
  int32_t out_col0;                    // FC: first output column of this op (a wide layer is cut into several ops)
  int32_t reserved;
};  // 17*4 + 2*4 + 4*8 + 9*4 + 2*128*2 + 2*4 = 664 bytes
#pragma pack(pop)
