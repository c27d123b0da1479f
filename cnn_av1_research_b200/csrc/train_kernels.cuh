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

// ---- data-parallel step: gradient reduce-scatter + AdamW + parameter all-gather in ONE kernel over NVLink peer memory.
//      Every rank owns one contiguous shard of the flat parameter index space (and only that shard's moments).  Per step
//      each rank launches this kernel after its backward; it
//        0. tells every peer "my gradients are complete" (one remote flag store per peer) and waits until all peers said so
//           - from here on nobody is still using the parameters or writing gradients;
//        1. for its shard: sums the W gradient buffers with peer loads in rank order, applies AdamW (adamw_one above, the
//           1 / W folded in), and stores the new parameter values into EVERY rank's parameter buffer with peer stores;
//        2. tells every peer "my shard is written everywhere and I am done reading your gradients" and its last CTA waits
//           for the same message from all peers, so that when the kernel completes on a rank its whole parameter buffer is
//           current and its gradient buffer is free for the next backward.
//      NVLink traffic per rank and step: (W - 1) / W of the gradients in, (W - 1) / W of the parameters out - half of an
//      all-reduce's, with no separate update pass over HBM afterwards.  Every shard is computed once, by one rank, in one
//      summation order, so the replicas are bit-identical by construction.  Flags live in each rank's own memory and are
//      written remotely / polled locally; they carry the step's epoch (monotonic), so they never need resetting.  Every spin
//      is bounded (~10 s of SM clocks): on a timeout the error word is set and the kernel leaves.
constexpr int DP_MAX_WORLD = 16;
constexpr int DP_FLAG_WORDS = 2 * DP_MAX_WORLD + 1;        // [0..15] phase-0 arrivals, [16..31] phase-2 arrivals, [32] finished CTAs
constexpr long long DP_SPIN_CYCLES = 20000000000ll;      // ~10 s

struct DpAdamWArgs {
  const float* grad[DP_MAX_WORLD];     // every rank's flat gradient buffer (peer-mapped), n elements
  float* param[DP_MAX_WORLD];          // every rank's flat parameter buffer
  int* flags[DP_MAX_WORLD];            // every rank's DP_FLAG_WORDS flag words
  float* m;                            // this rank's moment shards, indexed from the shard start
  float* v;
  long long n, shard;                  // total elements; elements per shard (multiple of 4)
  long long skip_lo, skip_hi;          // flat range without gradient (left untouched, as torch skips grad-less parameters)
  int rank, world, epoch;
  AdamWArgs h;                         // hyper-parameters, step pointer and grad_scale (p / g / m / v / n unused)
  int* err;
};

__device__ __forceinline__ void dp_store_flag(int* p, int v) { asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int dp_load_flag(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until every rank's word in my own flag array has reached `epoch`; false on timeout
__device__ __forceinline__ bool dp_wait_all(const int* mine, int world, int epoch) {
  const long long t0 = clock64();
  for (int p = 0; p < world; ++p)
    while (dp_load_flag(mine + p) - epoch < 0)
      if (clock64() - t0 > DP_SPIN_CYCLES) return false;
  return true;
}

__global__ void __launch_bounds__(TRAIN_THREADS) dp_adamw_fused_kernel(const DpAdamWArgs a) {
  __shared__ float s_step_size, s_bc2_sqrt;
  __shared__ int s_ok;
  int* mine = a.flags[a.rank];
  if (blockIdx.x == 0 && int(threadIdx.x) < a.world) {
    __threadfence_system();
    dp_store_flag(a.flags[threadIdx.x] + a.rank, a.epoch);            // phase 0: "rank's gradients are complete"
  }
  if (threadIdx.x == 0) {
    const double t = double(*a.h.step);
    s_step_size = float(a.h.lr / (1.0 - pow(a.h.beta1, t)));
    s_bc2_sqrt = float(sqrt(1.0 - pow(a.h.beta2, t)));
    s_ok = dp_wait_all(mine, a.world, a.epoch) ? 1 : 0;
    __threadfence_system();
  }
  __syncthreads();
  if (!s_ok) {
    if (threadIdx.x == 0) *a.err = 1;
    return;
  }
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const long long lo = min(a.n, (long long)a.rank * a.shard), hi = min(a.n, lo + a.shard);
  const long long n4 = (hi - lo) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long e = lo + (i << 2);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < a.world; ++p) {                                // fixed rank order: one summation order per element
      const float4 q = __ldcv(reinterpret_cast<const float4*>(a.grad[p] + e));
      g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
    }
    float4 w = *reinterpret_cast<const float4*>(a.param[a.rank] + e);
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    if (e + 3 < a.skip_lo || e >= a.skip_hi) {
      adamw_one(w.x, g.x, m.x, v.x, a.h, step_size, bc2_sqrt);
      adamw_one(w.y, g.y, m.y, v.y, a.h, step_size, bc2_sqrt);
      adamw_one(w.z, g.z, m.z, v.z, a.h, step_size, bc2_sqrt);
      adamw_one(w.w, g.w, m.w, v.w, a.h, step_size, bc2_sqrt);
    } else {                                                           // the quad overlaps the grad-less range
      if (e + 0 < a.skip_lo || e + 0 >= a.skip_hi) adamw_one(w.x, g.x, m.x, v.x, a.h, step_size, bc2_sqrt);
      if (e + 1 < a.skip_lo || e + 1 >= a.skip_hi) adamw_one(w.y, g.y, m.y, v.y, a.h, step_size, bc2_sqrt);
      if (e + 2 < a.skip_lo || e + 2 >= a.skip_hi) adamw_one(w.z, g.z, m.z, v.z, a.h, step_size, bc2_sqrt);
      if (e + 3 < a.skip_lo || e + 3 >= a.skip_hi) adamw_one(w.w, g.w, m.w, v.w, a.h, step_size, bc2_sqrt);
    }
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    for (int p = 0; p < a.world; ++p) *reinterpret_cast<float4*>(a.param[p] + e) = w;      // all-gather by peer stores
  }
  if (blockIdx.x == 0 && threadIdx.x < 3) {                            // scalar tail of the last shard (n not a multiple of 4)
    const long long e = lo + (n4 << 2) + threadIdx.x;
    if (e < hi) {
      float g = 0.f;
      for (int p = 0; p < a.world; ++p) g += __ldcv(a.grad[p] + e);
      float w = a.param[a.rank][e], m = a.m[e - lo], v = a.v[e - lo];
      if (e < a.skip_lo || e >= a.skip_hi) adamw_one(w, g, m, v, a.h, step_size, bc2_sqrt);
      a.m[e - lo] = m;
      a.v[e - lo] = v;
      for (int p = 0; p < a.world; ++p) a.param[p][e] = w;
    }
  }
  __threadfence_system();                                              // my peer stores are visible before anyone sees my flag
  __syncthreads();
  if (threadIdx.x == 0) {
    const int finished = atomicAdd(mine + 2 * DP_MAX_WORLD, 1);
    if (finished == int(gridDim.x) - 1) {                              // last CTA of this rank
      mine[2 * DP_MAX_WORLD] = 0;
      __threadfence_system();
      for (int p = 0; p < a.world; ++p) dp_store_flag(a.flags[p] + DP_MAX_WORLD + a.rank, a.epoch);   // phase 2
      if (!dp_wait_all(mine + DP_MAX_WORLD, a.world, a.epoch)) *a.err = 2;
      __threadfence_system();
    }
  }
}

}  // namespace av1p
