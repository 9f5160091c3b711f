// Fused optimiser step on flat fp32 arenas: global grad-norm, clip, adaptive-KL learning rate, Adam, refresh of
// the bf16 weight shadows the tensor-core layers read, and gradient zeroing -- two launches, no host sync.
// Replaces nn.utils.clip_grad_norm_ + optimizer.step + the lr schedule of reference ppo.py:162-174,211-214
// (~10 foreach kernels x 23 tensors + 3 host syncs per minibatch there).
// HBM traffic: 4 B/param (norm) + 16 B read + 12 B written (+2 B shadow) per parameter.
#include "hg_common.cuh"

namespace hg {

struct ShadowTable {
  HgBf16Shadow s[HG_MAX_SHADOWS];
  int n;
};

// One thread: everything the reference does on the host between backward() and optimizer.step().
__device__ void opt_scalars(const HgAdamArgs& a, double sumsq) {
  const float* stats = a.grads + a.count;
  const double rows = (double)a.rows_global;
  double lr = a.opt_state[0];
  const float kl_mean = (float)((double)stats[3] / rows);
  if (a.adaptive) {                                    // ppo.py:169-172
    if (kl_mean > a.desired_kl * 2.0f) lr = fmax((double)a.lr_min, lr / 1.5);
    else if (kl_mean < a.desired_kl / 2.0f && kl_mean > 0.0f) lr = fmin((double)a.lr_max, lr * 1.5);
  }
  const double step = a.opt_state[1] + 1.0;
  const double bc1 = 1.0 - pow((double)a.beta1, step);
  const double bc2 = 1.0 - pow((double)a.beta2, step);
  const float total = (float)sqrt(sumsq);               // clip_grad_norm_: 2-norm over all parameters
  float coef = a.max_grad_norm / (total + 1e-6f);
  coef = coef > 1.0f ? 1.0f : coef;
  a.opt_state[0] = lr;
  a.opt_state[1] = step;
  a.opt_state[2] = 0.0;
  a.opt_state[3] = (double)coef;
  a.opt_state[4] = lr / bc1;
  a.opt_state[5] = sqrt(bc2);
  a.opt_state[6] = (double)total;
  a.opt_state[7] = 0.0;                                 // block ticket counter of grad_sumsq_kernel
  if (a.loss_log) {
    double* log = a.loss_log + (size_t)a.log_slot * 8;
    log[0] = (double)stats[0] / rows;
    log[1] = (double)stats[1] / rows;
    log[2] = (double)stats[2] / (3.0 * rows);
    log[3] = (double)kl_mean;
    log[4] = (double)stats[4] / rows;
    log[5] = lr;
    log[6] = (double)total;
    log[7] = (double)coef;
  }
}

// Global gradient sum of squares (double accumulation); the LAST block to finish (ticket in opt_state[7]) derives the
// step's scalars -- clip coefficient, adaptive-KL learning rate, bias corrections -- so no separate one-thread launch.
__global__ void grad_sumsq_kernel(HgAdamArgs a) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();   // 592 x 256 threads: the whole grid is resident (see hg_common.cuh)
  const float* __restrict__ g = a.grads;
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += stride) {
    const float v = g[i];
    acc += (double)v * (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += s[w];
    atomicAdd(&a.opt_state[2], t);
    __threadfence();
    const double ticket = atomicAdd(&a.opt_state[7], 1.0);
    if (ticket == (double)(gridDim.x - 1)) {
      __threadfence();
      opt_scalars(a, atomicAdd(&a.opt_state[2], 0.0));   // every block's contribution is in (their adds precede their tickets)
    }
  }
}

// Four consecutive parameters per thread (tensor starts are 16-byte aligned and padded to multiples of four, so a float4
// never straddles two tensors): 16-byte loads / stores of p, g, m, v, and ONE search of the shadow table + one division per
// float4 instead of per element (the per-element version spent most of its 20 us there: ~11 table entries x 4 elements).
__global__ void __launch_bounds__(256) adam_kernel(HgAdamArgs a, const __grid_constant__ ShadowTable tab) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();
  const float coef = (float)a.opt_state[3];
  const float step_size = (float)a.opt_state[4];
  const float bc2_sqrt = (float)a.opt_state[5];
  const float w1 = 1.0f - a.beta1, w2 = 1.0f - a.beta2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long total4 = (a.count + HG_STATS_SLOTS) >> 2;      // count and HG_STATS_SLOTS are multiples of four
  const long long count4 = a.count >> 2;
  float4* __restrict__ P4 = reinterpret_cast<float4*>(a.params);
  float4* __restrict__ G4 = reinterpret_cast<float4*>(a.grads);
  float4* __restrict__ M4 = reinterpret_cast<float4*>(a.exp_avg);
  float4* __restrict__ V4 = reinterpret_cast<float4*>(a.exp_avg_sq);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += stride) {
    if (q >= count4) {                                  // statistics slots behind the gradients
      if (a.zero_grads) G4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float4 g4 = G4[q], m4 = M4[q], v4 = V4[q], p4 = P4[q];
    float g[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
    float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      m[e] = m[e] + w1 * (g[e] - m[e]);                 // exp_avg.lerp_(grad, 1-beta1)
      v[e] = v[e] * a.beta2 + w2 * g[e] * g[e];         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
      const float denom = sqrtf(v[e]) / bc2_sqrt + a.eps;
      p[e] = p[e] - step_size * (m[e] / denom);
    }
    P4[q] = make_float4(p[0], p[1], p[2], p[3]);
    M4[q] = make_float4(m[0], m[1], m[2], m[3]);
    V4[q] = make_float4(v[0], v[1], v[2], v[3]);
    if (a.zero_grads) G4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long i = q << 2;
    for (int s = 0; s < tab.n; ++s) {
      const long long off = i - tab.s[s].param_offset;
      if (off >= 0 && off < (long long)tab.s[s].rows * tab.s[s].cols) {
        const unsigned o32 = (unsigned)off, cols = (unsigned)tab.s[s].cols;      // a weight matrix has < 2^32 elements
        unsigned r = o32 / cols, c = o32 - r * cols;
        const unsigned sf = (unsigned)tab.s[s].src_frame, df = (unsigned)tab.s[s].dst_frame;
        unsigned f = 0, k = c;                           // frame-pitched copy (history-ring first layers): column f*sf + k -> f*df + k
        if (sf) { f = c / sf; k = c - f * sf; }
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(tab.s[s].dst);
#pragma unroll
        for (int e = 0; e < 4; ++e) {                    // (r, f, k) stepped along the four elements: one division for the lot
          if (o32 + e >= (unsigned)tab.s[s].rows * cols) break;      // (padding behind a tensor whose size is not a multiple of 4)
          dst[(long long)r * tab.s[s].ld + (sf ? f * df + k : c)] = __float2bfloat16_rn(p[e]);
          ++c; ++k;
          if (sf && k == sf) { k = 0; ++f; }
          if (c == cols) { c = 0; k = 0; f = 0; ++r; }
        }
        // no break: a first-layer weight has two shadows (packed and frame-pitched)
      }
    }
  }
}

}  // namespace hg

extern "C" int hg_clip_adam_step(const HgAdamArgs* a, void* stream) {
  using namespace hg;
  HG_CHECK(a && a->params && a->grads && a->exp_avg && a->exp_avg_sq && a->opt_state && a->count > 0,
           "hg_clip_adam_step: bad arguments");
  HG_CHECK(a->num_shadows >= 0 && a->num_shadows <= HG_MAX_SHADOWS, "at most %d bf16 shadows", HG_MAX_SHADOWS);
  HG_CHECK(a->num_shadows == 0 || a->shadows_host, "shadows_host is null");
  HG_CHECK(a->rows_global > 0, "rows_global must be positive");
  HG_CHECK(a->count % 4 == 0 && ((reinterpret_cast<uintptr_t>(a->params) | reinterpret_cast<uintptr_t>(a->grads) |
                                  reinterpret_cast<uintptr_t>(a->exp_avg) | reinterpret_cast<uintptr_t>(a->exp_avg_sq)) & 15) == 0,
           "hg_clip_adam_step: the arenas must be 16-byte aligned and count a multiple of 4 (tensors padded to 4 floats)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ShadowTable tab;
  memset(&tab, 0, sizeof(tab));
  tab.n = a->num_shadows;
  for (int i = 0; i < tab.n; ++i) tab.s[i] = a->shadows_host[i];
  HgAdamArgs dev = *a;
  dev.shadows_host = nullptr;
  int blocks = (int)((a->count + 256 * 4 - 1) / (256 * 4));
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  if (blocks < 1) blocks = 1;
  HG_CUDA(launch_kernel(grad_sumsq_kernel, dim3(blocks), dim3(256), 0, s, 1, dev));
  HG_LAUNCH_CHECK();
  HG_CUDA(launch_kernel(adam_kernel, dim3(blocks), dim3(256), 0, s, 1, dev, tab));
  HG_LAUNCH_CHECK();
  count_launch(2);
  return HG_OK;
}
