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

constexpr int STEM_NB = 8;                 // blocks per CTA pass
constexpr int STEM_THREADS = 32 * STEM_NB; // one thread per (block, conv column ox, pair of conv rows)
constexpr int STEM_WPAD = 52;              // 49 taps padded to 13 float4
constexpr int STEM_TILE_H = 22, STEM_TILE_W = 24;   // 16x16 block + 3-pixel zero halo (width padded)
constexpr int STEM_CH_PASS = 32;           // output channels per pass (bounds the fp32 conv tile in smem)
constexpr int STEM_CONV_LD = STEM_CH_PASS + 1;      // floats per conv-output row (+1 pad -> no bank conflicts)
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

__global__ void __launch_bounds__(STEM_THREADS, 2) stem_kernel(const StemParams p) {
  extern __shared__ __align__(16) uint8_t stem_smem[];
  float* w_s = reinterpret_cast<float*>(stem_smem);                       // [64][52]
  float* b_s = w_s + 64 * STEM_WPAD;                                      // [64]
  float* tile = b_s + 64;                                                 // [NB][22][24]
  float* conv = tile + STEM_NB * STEM_TILE_H * STEM_TILE_W;                // [NB][64 pos][33] fp32, one channel pass

  const int n = p.n_dev ? *p.n_dev : p.n;
  const int groups = (n + STEM_NB - 1) / STEM_NB;
  if (int(blockIdx.x) >= groups) return;

  for (int i = threadIdx.x; i < 64 * STEM_WPAD; i += STEM_THREADS) w_s[i] = p.w[i];
  if (threadIdx.x < 64) b_s[threadIdx.x] = p.b[threadIdx.x];
  for (int i = threadIdx.x; i < STEM_NB * STEM_TILE_H * STEM_TILE_W; i += STEM_THREADS) tile[i] = 0.f;
  __syncthreads();

  const int tb = threadIdx.x >> 5;       // block slot in this pass
  const int lane = threadIdx.x & 31;
  const int ox = lane & 7;               // conv output column
  const int oy = (lane >> 3) * 2;        // first of the two conv output rows this thread computes

  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    // ---- gather + normalise: thread handles 8 pixels: row = lane/2, cols (lane%2)*8..+7 of block tb
    {
      const int r = grp * STEM_NB + tb;
      const int py = lane >> 1, px0 = (lane & 1) * 8;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (r < n) {
        const int g = p.idx ? p.idx[r] : r;
        if (p.in.kind == 0) {
          const int f = g / p.in.blocks_per_frame;
          const int gb = g - f * p.in.blocks_per_frame;
          const int by = gb / p.in.blocks_x, bx = gb - by * p.in.blocks_x;
          const int y = by * 16 + py, x0 = bx * 16 + px0;
          if (y < p.in.height) {
            const uint16_t* src = p.in.frames + size_t(f) * p.in.frame_stride + size_t(y) * p.in.pitch + x0;
            if (x0 + 7 < p.in.width && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
              const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
              const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[2 * j] = float(u[j] & 0xFFFFu);
                v[2 * j + 1] = float(u[j] >> 16);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (x0 + j < p.in.width) v[j] = float(__ldg(src + j));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __fdiv_rn(v[j], 1023.0f);
          }
        } else {
          const float4* src = reinterpret_cast<const float4*>(p.in.images + size_t(g) * 256 + py * 16 + px0);
          const float4 a = __ldg(src), b = __ldg(src + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
          v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
      }
      float* t = tile + (tb * STEM_TILE_H + py + 3) * STEM_TILE_W + px0 + 3;
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = v[j];
    }
    __syncthreads();

    // ---- conv1: the thread keeps the 9x7 input patch of its two output rows in registers and walks
    //      the channels two at a time (4 accumulators; every weight fetched from smem feeds 2 FMAs)
    float patch[9][7];
    {
      const float* t = tile + (tb * STEM_TILE_H + 2 * oy) * STEM_TILE_W + 2 * ox;
#pragma unroll
      for (int y = 0; y < 9; ++y)
#pragma unroll
        for (int x = 0; x < 7; ++x) patch[y][x] = t[y * STEM_TILE_W + x];
    }
    for (int pass = 0; pass < 64 / STEM_CH_PASS; ++pass) {
      float* c0 = conv + (tb * 64 + oy * 8 + ox) * STEM_CONV_LD;     // conv row oy
      float* c1 = c0 + 8 * STEM_CONV_LD;                             // conv row oy + 1
#pragma unroll 1
      for (int cc = 0; cc < STEM_CH_PASS; cc += 2) {
        const int c = pass * STEM_CH_PASS + cc;
        float a00 = b_s[c], a01 = a00, a10 = b_s[c + 1], a11 = a10;  // a<channel><row>
        const float* w0 = w_s + c * STEM_WPAD;
        const float* w1 = w0 + STEM_WPAD;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            const float u0 = w0[ky * 7 + kx], u1 = w1[ky * 7 + kx];
            a00 = fmaf(patch[ky][kx], u0, a00);
            a01 = fmaf(patch[ky + 2][kx], u0, a01);
            a10 = fmaf(patch[ky][kx], u1, a10);
            a11 = fmaf(patch[ky + 2][kx], u1, a11);
          }
        }
        c0[cc] = fmaxf(a00, 0.f);
        c0[cc + 1] = fmaxf(a10, 0.f);
        c1[cc] = fmaxf(a01, 0.f);
        c1[cc + 1] = fmaxf(a11, 0.f);
      }
      __syncthreads();

      // ---- maxpool 3x3 s2 p1 (8x8 -> 4x4) in fp32, then fp16 (hi, and lo in split precision) stores
      for (int o = threadIdx.x; o < STEM_NB * 16 * (STEM_CH_PASS / 2); o += STEM_THREADS) {
        const int cp = o % (STEM_CH_PASS / 2);        // channel pair inside this pass
        const int q = (o / (STEM_CH_PASS / 2)) & 15;  // pooled position
        const int b = o / (16 * (STEM_CH_PASS / 2));
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
        const size_t off = size_t(r) * 1024 + q * 64 + pass * STEM_CH_PASS + 2 * cp;
        *reinterpret_cast<__half2*>(p.out + off) = hi;
        if (p.out_lo) {
          const float2 hf = __half22float2(hi);
          *reinterpret_cast<__half2*>(p.out_lo + off) = __floats2half2_rn(m0 - hf.x, m1 - hf.y);
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace av1p
