// Stem kernel: block gather (straight from the planar 10-bit frame, or from a float block
// tensor) -> /1023 normalisation -> conv1 7x7 s2 p3 (BN folded) -> ReLU -> maxpool 3x3 s2 p1
// -> fp16 activation rows [block][4x4 positions][64 channels].
//
// Reference semantics: pesquisa_v5/005_rearrange_video_YUV_420_10bit_LOSSLESS.py:353-457 (tiling,
// zero pad bottom/right, row-major block order), pesquisa_v6/v6_pipeline/data_hub.py:70-77
// (float32(u16) / 1023.0, true division) and models.py:105-108 (conv1/bn1/relu/maxpool).
// The convolution runs in fp32 on the CUDA cores so that the 10-bit input is consumed exactly;
// only the pooled output is rounded to fp16 for the tensor-core layers that follow.
#pragma once
#include "ptx_sm100.cuh"

namespace av1p {

constexpr int STEM_NB = 4;                 // blocks per CTA pass
constexpr int STEM_THREADS = 64 * STEM_NB; // one thread per (block, conv output position)
constexpr int STEM_WPAD = 52;              // 49 taps padded to 13 float4
constexpr int STEM_TILE_H = 22, STEM_TILE_W = 24;   // 16x16 block + 3-pixel zero halo (width padded)
constexpr int STEM_CONV_LD = 65;           // floats per conv-output row (64 ch + 1 pad -> no bank conflicts)
constexpr int STEM_SMEM_BYTES = 64 * STEM_WPAD * 4 + 64 * 4 + STEM_NB * STEM_TILE_H * STEM_TILE_W * 4 +
                                STEM_NB * 64 * STEM_CONV_LD * 4;

struct StemInput {
  // kind 0: planar YUV 4:2:0 10-bit LE frames resident in HBM.  Block id g -> frame g / blocks_per_frame,
  //         grid row (g % bpf) / blocks_x, grid col (g % bpf) % blocks_x.
  // kind 1: float32 blocks [n][16*16] (the tensor HierarchicalPipelineV6.predict receives).
  int kind;
  const uint16_t* frames;
  long long frame_stride;   // elements between consecutive frames (Y + U + V)
  int width, height, pitch; // luma geometry, pitch in elements
  int blocks_x, blocks_per_frame;
  const float* images;
};

struct StemParams {
  StemInput in;
  const int* idx;           // optional gather list: row r processes block id idx[r]
  const int* n_dev;         // device-side row count (nullptr -> n)
  int n;
  const float* w;           // [64][STEM_WPAD] folded conv1 weights (fp32), tap order ky*7+kx
  const float* b;           // [64] folded bias
  __half* out;              // [rows][1024]
  __half* out_lo;           // split precision: fp16(x - fp16(x)), nullptr otherwise
};

__global__ void __launch_bounds__(STEM_THREADS) stem_kernel(const StemParams p) {
  extern __shared__ __align__(16) uint8_t stem_smem[];
  float* w_s = reinterpret_cast<float*>(stem_smem);                       // [64][52]
  float* b_s = w_s + 64 * STEM_WPAD;                                      // [64]
  float* tile = b_s + 64;                                                 // [NB][22][24]
  float* conv = tile + STEM_NB * STEM_TILE_H * STEM_TILE_W;                // [NB][64][65] fp32

  const int n = p.n_dev ? *p.n_dev : p.n;
  const int groups = (n + STEM_NB - 1) / STEM_NB;
  if (int(blockIdx.x) >= groups) return;

  for (int i = threadIdx.x; i < 64 * STEM_WPAD; i += STEM_THREADS) w_s[i] = p.w[i];
  if (threadIdx.x < 64) b_s[threadIdx.x] = p.b[threadIdx.x];
  for (int i = threadIdx.x; i < STEM_NB * STEM_TILE_H * STEM_TILE_W; i += STEM_THREADS) tile[i] = 0.f;
  __syncthreads();

  const int tb = threadIdx.x >> 6;       // block slot in this pass
  const int pos = threadIdx.x & 63;      // conv output position 0..63 (8x8)
  const int oy = pos >> 3, ox = pos & 7;

  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    // ---- gather + normalise: thread handles 4 pixels: row = pos/4, cols (pos%4)*4..+3 of block tb
    {
      const int r = grp * STEM_NB + tb;
      const int py = pos >> 2, px0 = (pos & 3) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < n) {
        const int g = p.idx ? p.idx[r] : r;
        if (p.in.kind == 0) {
          const int f = g / p.in.blocks_per_frame;
          const int gb = g - f * p.in.blocks_per_frame;
          const int by = gb / p.in.blocks_x, bx = gb - by * p.in.blocks_x;
          const int y = by * 16 + py, x0 = bx * 16 + px0;
          if (y < p.in.height) {
            const uint16_t* src = p.in.frames + size_t(f) * p.in.frame_stride + size_t(y) * p.in.pitch + x0;
            if (x0 + 3 < p.in.width && ((reinterpret_cast<uintptr_t>(src) & 7u) == 0)) {
              const uint2 q = __ldg(reinterpret_cast<const uint2*>(src));
              v[0] = float(q.x & 0xFFFFu);
              v[1] = float(q.x >> 16);
              v[2] = float(q.y & 0xFFFFu);
              v[3] = float(q.y >> 16);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (x0 + j < p.in.width) v[j] = float(__ldg(src + j));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __fdiv_rn(v[j], 1023.0f);
          }
        } else {
          const float4 q = __ldg(reinterpret_cast<const float4*>(p.in.images + size_t(g) * 256 + py * 16 + px0));
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        }
      }
      float* t = tile + (tb * STEM_TILE_H + py + 3) * STEM_TILE_W + px0 + 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = v[j];
    }
    __syncthreads();

    // ---- conv1: thread (tb, pos) holds its 7x7 patch, loops over the 64 output channels
    float patch[49];
    {
      const float* t = tile + (tb * STEM_TILE_H + 2 * oy) * STEM_TILE_W + 2 * ox;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) patch[ky * 7 + kx] = t[ky * STEM_TILE_W + kx];
    }
    float* crow = conv + (tb * 64 + pos) * STEM_CONV_LD;
#pragma unroll 1
    for (int c = 0; c < 64; c += 2) {
      float a0 = b_s[c], a1 = b_s[c + 1];
      const float4* w0 = reinterpret_cast<const float4*>(w_s + c * STEM_WPAD);
      const float4* w1 = reinterpret_cast<const float4*>(w_s + (c + 1) * STEM_WPAD);
#pragma unroll
      for (int q = 0; q < 12; ++q) {
        const float4 u0 = w0[q], u1 = w1[q];
        a0 = fmaf(patch[4 * q + 0], u0.x, a0); a1 = fmaf(patch[4 * q + 0], u1.x, a1);
        a0 = fmaf(patch[4 * q + 1], u0.y, a0); a1 = fmaf(patch[4 * q + 1], u1.y, a1);
        a0 = fmaf(patch[4 * q + 2], u0.z, a0); a1 = fmaf(patch[4 * q + 2], u1.z, a1);
        a0 = fmaf(patch[4 * q + 3], u0.w, a0); a1 = fmaf(patch[4 * q + 3], u1.w, a1);
      }
      a0 = fmaf(patch[48], w_s[c * STEM_WPAD + 48], a0);
      a1 = fmaf(patch[48], w_s[(c + 1) * STEM_WPAD + 48], a1);
      crow[c] = fmaxf(a0, 0.f);
      crow[c + 1] = fmaxf(a1, 0.f);
    }
    __syncthreads();

    // ---- maxpool 3x3 s2 p1 (8x8 -> 4x4) in fp32, then fp16 (hi, and lo in split precision) stores
    for (int o = threadIdx.x; o < STEM_NB * 16 * 32; o += STEM_THREADS) {
      const int cp = o & 31;             // channel pair
      const int q = (o >> 5) & 15;       // pooled position
      const int b = o >> 9;
      const int r = grp * STEM_NB + b;
      if (r >= n) continue;
      const int qy = q >> 2, qx = q & 3;
      float m0 = 0.f, m1 = 0.f;          // inputs are post-ReLU (>= 0): 0 is the identity of max
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * qy + dy;
        if (y < 0 || y > 7) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = 2 * qx + dx;
          if (x < 0 || x > 7) continue;
          const float* cv = conv + (b * 64 + y * 8 + x) * STEM_CONV_LD + 2 * cp;
          m0 = fmaxf(m0, cv[0]);
          m1 = fmaxf(m1, cv[1]);
        }
      }
      const __half2 hi = __floats2half2_rn(m0, m1);
      const size_t off = size_t(r) * 1024 + q * 64 + 2 * cp;
      *reinterpret_cast<__half2*>(p.out + off) = hi;
      if (p.out_lo) {
        const float2 hf = __half22float2(hi);
        *reinterpret_cast<__half2*>(p.out_lo + off) = __floats2half2_rn(m0 - hf.x, m1 - hf.y);
      }
    }
    __syncthreads();
  }
}

}  // namespace av1p
