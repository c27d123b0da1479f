// Kernels of the Stage-1 data-parallel training step (BASELINE configs[4]) that are not
// convolutions: the focal loss with its gradient in ONE launch and the AdamW update of ALL parameters in ONE launch over
// flat fp32 buffers (parameters, averaged gradients and both moments live in one allocation each, in the order backward
// fills the gradient buffer - cnn_av1_research_b200/training.py).  The reference does both through ~15 elementwise
// PyTorch launches for the loss (pesquisa_v6/v6_pipeline/losses.py:29-38, :48-49) plus their autograd mirror images, and
// through torch.optim.AdamW's per-tensor / multi-tensor loops over 62 parameter tensors
// (pesquisa_v6/scripts/003_train_stage1_improved.py:64-73, 250-254).  Both kernels are HBM-bound by construction: the
// update streams 16 bytes in and 12 bytes out per parameter, once.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace av1p {

constexpr int TRAIN_THREADS = 256;

// ---- step counter: t <- t + 1 on the device, so that a captured CUDA graph of the step needs no host-side argument
__global__ void train_step_inc_kernel(int32_t* step) { *step += 1; }

// ---- AdamW (decoupled weight decay, Loshchilov & Hutter), torch.optim.AdamW semantics without amsgrad / maximize:
//        g  = grad * grad_scale                      (grad_scale = 1 / world size: the all-reduce SUMS)
//        p  = p * (1 - lr * wd)
//        m  = m + (g - m) * (1 - beta1)              (torch: exp_avg.lerp_(grad, 1 - beta1))
//        v  = v * beta2 + (1 - beta2) * g * g
//        p  = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
//      The hyper-parameters arrive as doubles (Python floats in the reference's call); 1 - beta, 1 - lr * wd and the bias
//      corrections are formed in double and rounded to fp32 once, as torch's scalar arguments are - `1.0f - 0.999f` differs
//      from fp32(1 - 0.999) by 1.3e-5 relative, which would show in exp_avg_sq.
//      t is read from device memory (already incremented for this step).  n may be any size; the four buffers must share
//      their address modulo 16 (4-byte aligned): scalar head up to the first 16-byte boundary, float4 body, scalar tail.
struct AdamWArgs {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  double lr, beta1, beta2;                       // for the bias corrections (evaluated in double, like torch does on the host)
  float decay, one_m_beta1, beta2_f, one_m_beta2, eps, grad_scale;      // rounded once from the double hyper-parameters
  const int32_t* step;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a, float step_size, float bc2_sqrt) {
  g *= a.grad_scale;
  p *= a.decay;
  m = fmaf(g - m, a.one_m_beta1, m);
  v = fmaf(v, a.beta2_f, a.one_m_beta2 * g * g);
  const float denom = sqrtf(v) / bc2_sqrt + a.eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(TRAIN_THREADS) adamw_flat_kernel(const AdamWArgs a) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = double(*a.step);
    const double bc1 = 1.0 - pow(a.beta1, t);
    const double bc2 = 1.0 - pow(a.beta2, t);
    s_step_size = float(a.lr / bc1);
    s_bc2_sqrt = float(sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  // scalar head up to the first 16-byte boundary (all four buffers share their alignment), float4 body, scalar tail
  const long long head = min((long long)(((16u - unsigned(reinterpret_cast<uintptr_t>(a.p) & 15u)) & 15u) >> 2), a.n);
  const long long n4 = (a.n - head) >> 2;
  const long long tail0 = head + (n4 << 2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(a.p + head);
  const float4* g4 = reinterpret_cast<const float4*>(a.g + head);
  float4* m4 = reinterpret_cast<float4*>(a.m + head);
  float4* v4 = reinterpret_cast<float4*>(a.v + head);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = p4[i];
    const float4 g = __ldcs(g4 + i);      // gradients are dead after this pass: streaming load
    float4 m = m4[i];
    float4 v = v4[i];
    adamw_one(p.x, g.x, m.x, v.x, a, step_size, bc2_sqrt);
    adamw_one(p.y, g.y, m.y, v.y, a, step_size, bc2_sqrt);
    adamw_one(p.z, g.z, m.z, v.z, a, step_size, bc2_sqrt);
    adamw_one(p.w, g.w, m.w, v.w, a, step_size, bc2_sqrt);
    p4[i] = p;
    m4[i] = m;
    v4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < 8) {
    // threads 0..3: head elements, threads 4..7: tail elements
    const long long i = threadIdx.x < 4 ? (long long)threadIdx.x : tail0 + (threadIdx.x - 4);
    const bool live = threadIdx.x < 4 ? i < head : i < a.n;
    if (live) {
      float p = a.p[i], m = a.m[i], v = a.v[i];
      adamw_one(p, a.g[i], m, v, a, step_size, bc2_sqrt);
      a.p[i] = p;
      a.m[i] = m;
      a.v[i] = v;
    }
  }
}

// ---- binary focal loss, mean reduction, with d(mean loss) / d(logit) (losses.py:29-38, 48-49):
//        z = x for target 1, -x for target 0;  pt = sigmoid(z);  bce = -log(pt);  a_t = alpha / (1 - alpha)
//        loss_i = a_t (1 - pt)^gamma bce
//        d loss_i / dx = sign * a_t * (1 - pt)^gamma * (gamma * pt * log(pt) - (1 - pt)),   sign = +1 / -1 for target 1 / 0
//      (from d pt / dz = pt (1 - pt)).  log(pt) = -softplus(-z) and 1 - pt = sigmoid(-z) are evaluated without cancellation.
//      One CTA: a training batch is 128 logits; the fixed-order tree reduction makes the loss bitwise reproducible.
constexpr int FOCAL_THREADS = 512;

__global__ void __launch_bounds__(FOCAL_THREADS) focal_loss_binary_kernel(const float* __restrict__ x, const long long* __restrict__ target, int n,
                                                                           float alpha, float gamma, float* __restrict__ loss_out,
                                                                           float* __restrict__ dx) {
  __shared__ float s_sum[FOCAL_THREADS];
  float acc = 0.f;
  const float inv_n = 1.0f / float(n);
  for (int i = threadIdx.x; i < n; i += FOCAL_THREADS) {
    const bool pos = target[i] != 0;
    const float z = pos ? x[i] : -x[i];
    // log(pt) = -log(1 + exp(-z)); exp argument is never positive
    const float e = expf(-fabsf(z));
    const float log_pt = (z >= 0.f ? 0.f : z) - log1pf(e);
    const float pt = z >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
    const float one_m_pt = z >= 0.f ? e / (1.0f + e) : 1.0f / (1.0f + e);
    const float a_t = pos ? alpha : 1.0f - alpha;
    const float w = a_t * powf(one_m_pt, gamma);
    acc += w * (-log_pt);
    if (dx) dx[i] = (pos ? 1.0f : -1.0f) * w * (gamma * pt * log_pt - one_m_pt) * inv_n;
  }
  s_sum[threadIdx.x] = acc;
  __syncthreads();
#pragma unroll
  for (int s = FOCAL_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_sum[threadIdx.x] += s_sum[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = s_sum[0] * inv_n;
}

}  // namespace av1p
