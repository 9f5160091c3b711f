// Output layers ("heads") of the three MLPs, in fp32 SIMT, fused with everything that follows them.
//
// The last Linear of each net is tiny (H x {A,1,3}, H = 128 for XBot) and feeds scalar-heavy math, so it
// stays in fp32 on CUDA cores and is fused:
//   rollout : mu = W_a h_a + b_a, a = mu + std*eps, log-prob, V = W_c h_c + b_c     (PPO.act, ppo.py:116-126)
//   update  : forward of the 3 heads + log-prob/KL/ratio/clipped surrogate/clipped value loss/entropy/lin-vel
//             MSE (ppo.py:154-208) + dLoss/dOutputs + backward through the output Linear (dW, db, dH) and
//             through the ELU of the last hidden layer (dZ, bias gradient), + d/d std.
// Keeping the output layer in fp32 also keeps the policy mean's response to small Adam steps exact,
// which the PPO ratio is sensitive to ("numerics").
#include <stdlib.h>

#include "hg_common.cuh"

namespace hg {

constexpr float HALF_LOG_2PI_PLUS_HALF = 1.4189385332046727f;

// ------------------------------------------------------------------------------------------ shared helpers
// Stage `rows` rows of a bf16 activation matrix into shared memory as [rows][H + 8] (the +8 keeps 16-byte row
// alignment and staggers banks), rows beyond `valid` are zero.
__device__ __forceinline__ void stage_h(const __nv_bfloat16* __restrict__ h, long long ld, int H, int row0, int valid,
                                        int rows, __nv_bfloat16* sh) {
  const int chunks = H >> 3;
  const int pitch = H + 8;
  const int total = rows * chunks;
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {      // 4 independent 16-byte loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = make_uint4(0, 0, 0, 0);
      if (i < total) {
        const int r = i / chunks, c = i - r * chunks;
        if (r < valid) v[u] = __ldg(reinterpret_cast<const uint4*>(h + (long long)(row0 + r) * ld) + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total) {
        const int r = i / chunks, c = i - r * chunks;
        *reinterpret_cast<uint4*>(sh + r * pitch + c * 8) = v[u];
      }
    }
  }
}
__device__ __forceinline__ void stage_f32(const float* __restrict__ src, int n, float* dst) {
  // 16-byte loads, all issued before the first store (independent -> one memory round trip), scalar tail
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * blockDim.x < n4) v[u] = __ldg(reinterpret_cast<const float4*>(src) + i0 + u * blockDim.x);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * blockDim.x < n4) reinterpret_cast<float4*>(dst)[i0 + u * blockDim.x] = v[u];
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
  }
}

// Partial dot products of one staged row with n_out weight rows, for the k-chunks owned by `sub` of `nsub`
// cooperating lanes; reduced over the nsub lanes (nsub is a power of two <= 32, lanes are adjacent).
template <int MAX_OUT, bool EXACT = false>
__device__ __forceinline__ void head_forward_row(const __nv_bfloat16* hrow, int H, const float* W, const float* b,
                                                 int n_out, int sub, int nsub, float (&out)[MAX_OUT]) {
  // EXACT: n_out == MAX_OUT is known at compile time, so no per-output predication is generated
#pragma unroll
  for (int a = 0; a < MAX_OUT; ++a) out[a] = 0.f;
  for (int c = sub; c < (H >> 3); c += nsub) {
    const uint4 hv = *reinterpret_cast<const uint4*>(hrow + c * 8);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
    float hf[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(hw[k]);
      hf[2 * k] = f.x;
      hf[2 * k + 1] = f.y;
    }
#pragma unroll
    for (int a = 0; a < MAX_OUT; ++a) {
      if (EXACT || a < n_out) {
        const float4 w0 = *reinterpret_cast<const float4*>(W + a * H + c * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(W + a * H + c * 8 + 4);
        out[a] += hf[0] * w0.x + hf[1] * w0.y + hf[2] * w0.z + hf[3] * w0.w + hf[4] * w1.x + hf[5] * w1.y +
                  hf[6] * w1.z + hf[7] * w1.w;
      }
    }
  }
#pragma unroll
  for (int a = 0; a < MAX_OUT; ++a) {
    if (EXACT || a < n_out) {
      for (int o = nsub >> 1; o > 0; o >>= 1) out[a] += __shfl_xor_sync(0xffffffffu, out[a], o);
      out[a] += b[a];
    }
  }
}

// ------------------------------------------------------------------------------------------ rollout head
constexpr int RH_ROWS = 16;      // rows per block (256 blocks at 4096 envs)
constexpr int RH_SUB = 8;        // lanes per row
constexpr int MAX_A = 32;        // max num_actions (kernels are instantiated for action capacities 12 / 20 / 32)

template <int AT>
__global__ void __launch_bounds__(RH_ROWS* RH_SUB)
    policy_head_rollout_kernel(const __nv_bfloat16* __restrict__ ha, long long ld_ha, int H_a, const float* __restrict__ W_a,
                               const float* __restrict__ b_a, const float* __restrict__ std, const __nv_bfloat16* __restrict__ hc,
                               long long ld_hc, int H_c, const float* __restrict__ W_c, const float* __restrict__ b_c,
                               const float* __restrict__ eps, uint64_t seed, uint64_t counter, int n, int A,
                               float* __restrict__ actions, float* __restrict__ mu, float* __restrict__ sigma,
                               float* __restrict__ log_prob, float* __restrict__ values) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  float* sWa = reinterpret_cast<float*>(smem);                 // [A][H_a]
  float* sWc = sWa + A * H_a;                                  // [H_c]
  float* sb = sWc + H_c;                                       // [A] b_a, [1] b_c, [A] std
  __nv_bfloat16* sha = reinterpret_cast<__nv_bfloat16*>(sb + 2 * MAX_A + 4);
  __nv_bfloat16* shc = ha ? sha + RH_ROWS * (H_a + 8) : sha;
  const int row0 = blockIdx.x * RH_ROWS;
  const int valid = min(RH_ROWS, n - row0);
  if (ha) {
    stage_f32(W_a, A * H_a, sWa);
    stage_f32(b_a, A, sb);
    stage_f32(std, A, sb + MAX_A + 1);
    stage_h(ha, ld_ha, H_a, row0, valid, RH_ROWS, sha);
  }
  stage_f32(W_c, H_c, sWc);
  if (threadIdx.x == 0) sb[MAX_A] = b_c[0];
  stage_h(hc, ld_hc, H_c, row0, valid, RH_ROWS, shc);
  __syncthreads();

  const int r = threadIdx.x / RH_SUB, sub = threadIdx.x % RH_SUB;
  const int row = row0 + r;
  float v1[1];
  head_forward_row<1>(shc + r * (H_c + 8), H_c, sWc, sb + MAX_A, 1, sub, RH_SUB, v1);
  if (sub == 0 && row < n) values[row] = v1[0];
  if (!ha) return;
  float m[AT];
  head_forward_row<AT>(sha + r * (H_a + 8), H_a, sWa, sb, A, sub, RH_SUB, m);
  // the row's lanes split the actions; the log-prob terms are then summed over the lanes
  float lp = 0.f;
#pragma unroll
  for (int a = 0; a < AT; ++a) {
    if (a < A && (a % RH_SUB) == sub && row < n) {
      const float sd = sb[MAX_A + 1 + a];
      const float e = eps ? eps[(long long)row * A + a] : philox_normal(seed, counter, (uint32_t)row, (uint32_t)a);
      const float act = m[a] + sd * e;                       // Normal(mu, std).sample()  (actor_critic.py:129,133)
      const float d = act - m[a];
      lp += -(d * d) / (2.f * sd * sd) - logf(sd) - LOG_SQRT_2PI;   // Normal.log_prob
      actions[(long long)row * A + a] = act;
      mu[(long long)row * A + a] = m[a];
      sigma[(long long)row * A + a] = sd;
    }
  }
#pragma unroll
  for (int o = RH_SUB >> 1; o > 0; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
  if (sub == 0 && row < n) log_prob[row] = lp;
}

__global__ void __launch_bounds__(RH_ROWS* RH_SUB)
    head_forward_kernel(const __nv_bfloat16* __restrict__ h, long long ld_h, int H, const float* __restrict__ W,
                        const float* __restrict__ b, int n_out, int n, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  float* sW = reinterpret_cast<float*>(smem);                  // [n_out][H]
  float* sb = sW + n_out * H;                                  // [MAX_A]
  __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(sb + MAX_A);
  const int row0 = blockIdx.x * RH_ROWS;
  const int valid = min(RH_ROWS, n - row0);
  stage_f32(W, n_out * H, sW);
  stage_f32(b, n_out, sb);
  stage_h(h, ld_h, H, row0, valid, RH_ROWS, sh);
  __syncthreads();
  const int r = threadIdx.x / RH_SUB, sub = threadIdx.x % RH_SUB;
  float m[MAX_A];
  head_forward_row<MAX_A>(sh + r * (H + 8), H, sW, sb, n_out, sub, RH_SUB, m);
  if (row0 + r < n) {
#pragma unroll
    for (int a = 0; a < MAX_A; ++a)
      if (a < n_out && (a % RH_SUB) == sub) out[(long long)(row0 + r) * n_out + a] = m[a];
  }
}

// ------------------------------------------------------------------------------------------ update head + losses
constexpr int LH_ROWS = 64;
constexpr int LH_THREADS = 256;
constexpr int LH_SUB = LH_THREADS / LH_ROWS;   // 4 lanes per row in the forward / loss phase
constexpr int LH_GROUPS = LH_THREADS / 128;    // row groups in the backward phase

struct NetSmem {
  __nv_bfloat16* h;   // [LH_ROWS][H+8]
  float* W;           // [n_out][H]
  float* b;           // [n_out]
  float* dOut;        // [LH_ROWS][n_out]
};

// Backward of one head for the block's rows: column-parallel over the hidden index k.
//   dH[r,k] = sum_a dOut[r,a] W[a,k];  dZ = dH * elu'(h);  gW[a,k] += dOut[r,a] h[r,k];  gb_hidden[k] += dZ[r,k]
template <int NO, bool EXACT>
__device__ __forceinline__ void head_backward(const HgHeadNet& net, const NetSmem& s, int row0, int valid, float* red) {
  const int H = net.H, n_out = EXACT ? NO : net.n_out, pitch = H + 8;
  const int grp = threadIdx.x / 128;
  for (int k0 = 0; k0 < H; k0 += 128) {            // uniform trip count: the loop body holds block barriers
    const int k = k0 + (threadIdx.x % 128);
    const bool k_ok = k < H;
    float gw[NO], w[NO];
#pragma unroll
    for (int a = 0; a < NO; ++a) {
      gw[a] = 0.f;
      w[a] = (k_ok && (EXACT || a < n_out)) ? s.W[a * H + k] : 0.f;
    }
    float gbh = 0.f;
    if (k_ok) {
      for (int r = grp; r < valid; r += LH_GROUPS) {
        const float h = __bfloat162float(s.h[r * pitch + k]);
        float dd[NO];
        if constexpr (EXACT && (NO % 4 == 0)) {      // the row of output gradients as broadcast 16-byte loads
#pragma unroll
          for (int a = 0; a < NO; a += 4) {
            const float4 v = *reinterpret_cast<const float4*>(s.dOut + r * NO + a);
            dd[a] = v.x; dd[a + 1] = v.y; dd[a + 2] = v.z; dd[a + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int a = 0; a < NO; ++a) dd[a] = (EXACT || a < n_out) ? s.dOut[r * n_out + a] : 0.f;
        }
        float dh = 0.f;
#pragma unroll
        for (int a = 0; a < NO; ++a) {
          dh += dd[a] * w[a];
          gw[a] += dd[a] * h;
        }
        const float dz = dh * (h > 0.f ? 1.f : h + 1.f);
        reinterpret_cast<__nv_bfloat16*>(net.dZ)[(long long)(row0 + r) * net.ld_dz + k] = __float2bfloat16_rn(dz);
        gbh += dz;
      }
    }
    // combine the row groups through shared memory, then one atomic per (a,k) per block
    if (grp > 0) {
      float* slot = red + (size_t)((grp - 1) * 128 + (threadIdx.x % 128)) * (NO + 1);
#pragma unroll
      for (int a = 0; a < NO; ++a) slot[a] = gw[a];
      slot[NO] = gbh;
    }
    __syncthreads();
    if (grp == 0 && k_ok) {
      for (int g = 1; g < LH_GROUPS; ++g) {
        const float* slot = red + (size_t)((g - 1) * 128 + (threadIdx.x % 128)) * (NO + 1);
#pragma unroll
        for (int a = 0; a < NO; ++a) gw[a] += slot[a];
        gbh += slot[NO];
      }
#pragma unroll
      for (int a = 0; a < NO; ++a)
        if (EXACT || a < n_out) atomicAdd(net.gW + a * H + k, gw[a]);
      atomicAdd(net.gb_hidden + k, gbh);
    }
    __syncthreads();
  }
}

template <int AT, bool EXACT>
__global__ void __launch_bounds__(LH_THREADS) ppo_head_loss_kernel(const __grid_constant__ HgLossArgs L) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  const int A = EXACT ? AT : L.A;
  const int RS = 3 * A + 8;            // per-row gathered record: actions[A] old_mu[A] old_sigma[A] logp adv ret v tgt[3] pad
  // ---- carve shared memory
  uint8_t* p = smem;
  auto carve = [&](size_t bytes) { uint8_t* q = p; p += (bytes + 15) & ~size_t(15); return q; };
  NetSmem sa, sc, sl;
  const HgHeadNet* nets[3] = {&L.actor, &L.critic, &L.linvel};
  NetSmem* ss[3] = {&sa, &sc, &sl};
  for (int i = 0; i < 3; ++i) {
    ss[i]->h = reinterpret_cast<__nv_bfloat16*>(carve((size_t)LH_ROWS * (nets[i]->H + 8) * 2));
    ss[i]->W = reinterpret_cast<float*>(carve((size_t)nets[i]->n_out * nets[i]->H * 4));
    ss[i]->b = reinterpret_cast<float*>(carve((size_t)nets[i]->n_out * 4));
    ss[i]->dOut = reinterpret_cast<float*>(carve((size_t)LH_ROWS * nets[i]->n_out * 4));
  }
  float* s_std = reinterpret_cast<float*>(carve(MAX_A * 4));
  float* s_c = reinterpret_cast<float*>(carve(3 * MAX_A * 4));     // [1/(2 sigma^2) | log sigma | 1/sigma] per action
  float* s_rec = reinterpret_cast<float*>(carve((size_t)LH_ROWS * RS * 4));
  float* s_gs = reinterpret_cast<float*>(carve((size_t)LH_ROWS * A * 4));          // per-row d loss / d std[a]
  float* s_stat = reinterpret_cast<float*>(carve((size_t)(LH_THREADS / 32) * 8 * 4));
  long long* s_idx = reinterpret_cast<long long*>(carve((size_t)LH_ROWS * 8));
  float* s_red = reinterpret_cast<float*>(carve((size_t)(LH_GROUPS - 1) * 128 * (AT + 1) * 4 + 16));

  const int row0 = blockIdx.x * LH_ROWS;
  const int valid = min(LH_ROWS, L.rows - row0);
  if (threadIdx.x < LH_ROWS)
    s_idx[threadIdx.x] = threadIdx.x < valid ? (L.idx ? L.idx[row0 + threadIdx.x] : (long long)(row0 + threadIdx.x)) : 0;
  for (int i = 0; i < 3; ++i) {
    stage_h(reinterpret_cast<const __nv_bfloat16*>(nets[i]->h), nets[i]->ld_h, nets[i]->H, row0, valid, LH_ROWS, ss[i]->h);
    stage_f32(nets[i]->W, nets[i]->n_out * nets[i]->H, ss[i]->W);
    stage_f32(nets[i]->b, nets[i]->n_out, ss[i]->b);
  }
  stage_f32(L.std, A, s_std);
  if (threadIdx.x < A) {
    const float sd = __ldg(L.std + threadIdx.x);
    s_c[threadIdx.x] = 1.f / (2.f * sd * sd);
    s_c[MAX_A + threadIdx.x] = logf(sd);
    s_c[2 * MAX_A + threadIdx.x] = 1.f / sd;
  }
  __syncthreads();
  // ---- cooperative gather of the rows' old-policy records (all loads independent -> full memory-level parallelism)
  for (int i = threadIdx.x; i < valid * RS; i += LH_THREADS) {
    const int r = i / RS, c = i - r * RS;
    const long long src = s_idx[r];
    float v = 0.f;
    if (c < A) v = __ldg(L.actions + src * A + c);
    else if (c < 2 * A) v = __ldg(L.old_mu + src * A + (c - A));
    else if (c < 3 * A) v = __ldg(L.old_sigma + src * A + (c - 2 * A));
    else if (c == 3 * A) v = __ldg(L.old_log_prob + src);
    else if (c == 3 * A + 1) v = __ldg(L.advantages + src);
    else if (c == 3 * A + 2) v = __ldg(L.returns + src);
    else if (c == 3 * A + 3) v = __ldg(L.old_values + src);
    else if (c < 3 * A + 7) v = __ldg(L.critic_obs + src * L.ld_cobs + L.lv_col + (c - 3 * A - 4));
    s_rec[i] = v;
  }
  __syncthreads();

  // ---- forward of the three heads + losses, LH_SUB lanes per row
  const int r = threadIdx.x / LH_SUB, sub = threadIdx.x % LH_SUB;
  float m[AT], v1[1], lv[3];
  head_forward_row<AT, EXACT>(sa.h + r * (L.actor.H + 8), L.actor.H, sa.W, sa.b, A, sub, LH_SUB, m);
  head_forward_row<1, true>(sc.h + r * (L.critic.H + 8), L.critic.H, sc.W, sc.b, 1, sub, LH_SUB, v1);
  head_forward_row<3, true>(sl.h + r * (L.linvel.H + 8), L.linvel.H, sl.W, sl.b, 3, sub, LH_SUB, lv);

  float st_v = 0.f, st_s = 0.f, st_l = 0.f, st_k = 0.f, st_e = 0.f;
  {
    // The row's LH_SUB lanes split the actions (a = sub, sub+4, ...): Gaussian log-prob, entropy, KL(old || new)
    // (actor_critic.py:123-138, ppo.py:165-166), then a 2-step shuffle reduction; lane 0 of the row finishes the
    // scalar losses.  Per-action constants (1/(2 sigma^2), log sigma, ...) come from shared memory: no divisions
    // by sigma in the per-row path.
    const bool row_ok = r < valid;
    const float* rec = s_rec + r * RS;
    const float inv_rows = L.inv_world / (float)L.rows;      // 1 / (rows over all ranks)
    float logp = 0.f, ent = 0.f, kl = 0.f;
#pragma unroll
    for (int a = 0; a < AT; ++a) {
      if ((EXACT || a < A) && (a % LH_SUB) == sub && row_ok) {
        const float sd = s_std[a], i2v = s_c[a], lsd = s_c[MAX_A + a];
        const float d = rec[a] - m[a];
        logp += -(d * d) * i2v - lsd - LOG_SQRT_2PI;
        ent += HALF_LOG_2PI_PLUS_HALF + lsd;
        const float osd = rec[2 * A + a];
        const float dm = rec[A + a] - m[a];
        kl += logf(sd * __frcp_rn(osd) + 1.e-5f) + (osd * osd + dm * dm) * i2v - 0.5f;
      }
    }
#pragma unroll
    for (int o = LH_SUB >> 1; o > 0; o >>= 1) {
      logp += __shfl_xor_sync(0xffffffffu, logp, o);
      ent += __shfl_xor_sync(0xffffffffu, ent, o);
      kl += __shfl_xor_sync(0xffffffffu, kl, o);
    }
    float g_logp = 0.f;
    if (row_ok) {
      // -- clipped surrogate (ppo.py:178-184); every lane of the row computes it (needed for its actions' gradients)
      const float adv = rec[3 * A + 1];
      const float ratio = expf(logp - rec[3 * A]);
      const float s1 = -adv * ratio;
      const float s2 = -adv * fminf(fmaxf(ratio, 1.f - L.clip_param), 1.f + L.clip_param);
      g_logp = (s1 >= s2 ? s1 : 0.f) * inv_rows;             // d surr / d logp  (ties: clamp passes gradient)
      if (sub == 0) {
        st_s = fmaxf(s1, s2); st_k = kl; st_e = ent;
        // -- value loss (ppo.py:187-194)
        const float R = rec[3 * A + 2], v = v1[0];
        float vl, g_v;
        if (L.use_clipped_value_loss) {
          const float vo = rec[3 * A + 3];
          const float dv = v - vo;
          const float vc = vo + fminf(fmaxf(dv, -L.clip_param), L.clip_param);
          const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
          const float pass = (dv >= -L.clip_param && dv <= L.clip_param) ? 1.f : 0.f;
          vl = fmaxf(l1, l2);
          g_v = l1 > l2 ? 2.f * (v - R) : (l1 < l2 ? 2.f * (vc - R) * pass : (v - R) + (vc - R) * pass);
        } else {
          vl = (R - v) * (R - v);
          g_v = 2.f * (v - R);
        }
        st_v = vl;
        sc.dOut[r] = L.value_loss_coef * g_v * inv_rows;
        // -- lin-vel regression (ppo.py:205; target = critic_obs[:, 53:56], rollout_storage.py:164)
        float lsq = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float e = lv[j] - rec[3 * A + 4 + j];
          lsq += e * e;
          sl.dOut[r * 3 + j] = L.lin_vel_coef * 2.f * e * inv_rows * (1.f / 3.f);
        }
        st_l = lsq;
      }
    } else if (sub == 0) {
      sc.dOut[r] = 0.f;
      for (int j = 0; j < 3; ++j) sl.dOut[r * 3 + j] = 0.f;
    }
    // -- gradients wrt mean and std for this lane's actions
#pragma unroll
    for (int a = 0; a < AT; ++a) {
      if ((EXACT || a < A) && (a % LH_SUB) == sub) {
        float gm = 0.f, gs = 0.f;
        if (row_ok) {
          const float isd = s_c[2 * MAX_A + a], ivar = 2.f * s_c[a];
          const float d = rec[a] - m[a];
          gm = g_logp * d * ivar;
          gs = g_logp * (d * d * ivar * isd - isd) - L.entropy_coef * inv_rows * isd;
        }
        sa.dOut[r * A + a] = gm;
        s_gs[r * A + a] = gs;
      }
    }
  }
  // statistics: warp shuffle reduction, then one partial per warp in shared memory
  st_v = warp_sum(st_v); st_s = warp_sum(st_s); st_l = warp_sum(st_l); st_k = warp_sum(st_k); st_e = warp_sum(st_e);
  if ((threadIdx.x & 31) == 0) {
    float* w = s_stat + (threadIdx.x >> 5) * 8;
    w[0] = st_v; w[1] = st_s; w[2] = st_l; w[3] = st_k; w[4] = st_e;
  }
  __syncthreads();

  // ---- output-layer bias gradients, std gradient and statistics: a handful of atomics per block
  {
    const int t = threadIdx.x;
    if (t < A) {                                  // actor output bias + std
      float g = 0.f, gs = 0.f;
      for (int rr = 0; rr < LH_ROWS; ++rr) { g += sa.dOut[rr * A + t]; gs += s_gs[rr * A + t]; }
      atomicAdd(L.actor.gb + t, g);
      atomicAdd(L.g_std + t, gs);
    } else if (t >= 32 && t < 33) {
      float g = 0.f;
      for (int rr = 0; rr < LH_ROWS; ++rr) g += sc.dOut[rr];
      atomicAdd(L.critic.gb, g);
    } else if (t >= 64 && t < 67) {
      float g = 0.f;
      for (int rr = 0; rr < LH_ROWS; ++rr) g += sl.dOut[rr * 3 + (t - 64)];
      atomicAdd(L.linvel.gb + (t - 64), g);
    } else if (t >= 96 && t < 101) {
      float g = 0.f;
      for (int w = 0; w < LH_THREADS / 32; ++w) g += s_stat[w * 8 + (t - 96)];
      atomicAdd(L.stats + (t - 96), g);
    }
  }

  // ---- backward through the three output layers and the ELU below them
  head_backward<AT, EXACT>(L.actor, sa, row0, valid, s_red);
  head_backward<1, true>(L.critic, sc, row0, valid, s_red);
  head_backward<3, true>(L.linvel, sl, row0, valid, s_red);
}

// ------------------------------------------------------------------------------------------ update head on tensor cores
// Same operator as ppo_head_loss_kernel for the common shape (three heads of equal hidden width H <= 128, H % 32 == 0,
// A <= 16, A % 4 == 0, 16-byte aligned arrays): the three tiny GEMMs per row tile run on the tensor cores (warp-level
// mma.sync m16n8k16 bf16 -> fp32, operands through ldmatrix) and the blocks are PERSISTENT, so the weight gradients live
// in accumulator registers for the whole launch and are added to global memory once per block instead of once per tile.
// The tiles (32 rows) run through a two-stage software pipeline: while one is computed on, the next one's activation
// rows (TMA bulk copies, completion on an mbarrier) and old-policy record (cp.async) are in flight into the other buffer.
//   forward : out = h W^T with W = W_hi + W_mid + W_lo (three bf16 terms = all 24 mantissa bits; h is bf16, every
//             product is exact in fp32), so the means match the fp32 rollout head to accumulation-order rounding --
//             the PPO ratio is sensitive to exactly that.
//   backward: dH = dOut W and gW = dOut^T h with dOut = hi + lo (16 mantissa bits), W = hi + mid.
// The critic (1 output) and lin-vel (3 outputs) heads share one 16-row weight tile: row 0 = critic, rows 1..3 = lin-vel.
namespace mh {
constexpr int ROWS = 32;             // rows per tile
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int SUB = THREADS / ROWS;  // lanes per row in the loss phase
static_assert(SUB >= 7, "the SUB lanes of a row fetch its 7 scalars, one each");
constexpr int NO = 16;               // padded outputs per head (two n8 mma tiles)
constexpr int DP = NO + 8;           // row pitch of the gradient tiles (48 B: conflict-free ldmatrix)
constexpr int MAXP = 3;              // (net, 16-column tile) pairs per warp: 3 nets * (128/16) / 8 warps
constexpr int HMAX = 128;
constexpr int AJ = NO / SUB;         // actions per lane
// A row's old-policy record in shared memory: actions | old means | old sigmas (NO words each) | old log-prob, advantage,
// return, old value, 3 lin-vel targets.  56 words: 16-byte aligned rows, and the 4 rows a warp reads sit 24 banks apart.
constexpr int RS = 3 * NO + 8;
constexpr int RS_SCALARS = 3 * NO;

struct Layout {
  size_t sh, wm, out, dO, bias, cst, rec, bar, sh_buf, total;
  __host__ __device__ explicit Layout(int H) {
    const size_t HP = H + 8;
    size_t o = 0;
    auto take = [&](size_t b) { size_t q = o; o += (b + 127) & ~size_t(127); return q; };
    sh_buf = (3 * ROWS * HP * 2 + 127) & ~size_t(127);
    sh = take(2 * sh_buf);                       // two tiles: the next one lands while this one is computed on
    rec = take(2 * ROWS * RS * 4);
    wm = take(2 * 3 * NO * HP * 2);
    out = take(3 * ROWS * NO * 4);
    dO = take(3 * 2 * ROWS * DP * 2);
    bias = take(3 * NO * 4);
    cst = take(4 * NO * 4);
    bar = take(2 * 8);                           // one mbarrier per activation buffer
    total = o;
  }
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
// D[16x8] += A[16x16] B[16x8]  (bf16 operands, fp32 accumulate)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
// One contiguous run of bytes global -> shared through the TMA unit (no L1 miss slots, no per-thread address work);
// completion is counted on an mbarrier.  dst, src and bytes are multiples of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t saddr, const void* g, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(saddr), "l"(g), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool valid) {
  const int sz = valid ? 16 : 0;      // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
}  // namespace mh

__global__ void __launch_bounds__(mh::THREADS, 2) ppo_head_loss_mma_kernel(const __grid_constant__ HgLossArgs L) {
  using namespace mh;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) trace(0x000);
  extern __shared__ __align__(128) uint8_t smem[];
  const int H = L.actor.H, HP = H + 8, A = L.A, NT = H >> 4, P = 3 * NT;
  const Layout lay(H);
  __nv_bfloat16* sh0 = reinterpret_cast<__nv_bfloat16*>(smem + lay.sh);         // [2 buffers][3][ROWS][HP]
  const float* s_rec = reinterpret_cast<const float*>(smem + lay.rec);          // [2][ROWS][RS]
  __nv_bfloat16* wm = reinterpret_cast<__nv_bfloat16*>(smem + lay.wm);          // [2][3][NO][HP]
  float* s_out = reinterpret_cast<float*>(smem + lay.out);                      // [3][ROWS][NO]
  __nv_bfloat16* s_dO = reinterpret_cast<__nv_bfloat16*>(smem + lay.dO);        // [3][2][ROWS][DP]
  float* s_bias = reinterpret_cast<float*>(smem + lay.bias);                    // [3][NO]
  float* s_c = reinterpret_cast<float*>(smem + lay.cst);                        // [std | 1/(2 s^2) | log s | 1/s][NO]
  const uint32_t sh0_u = smem_u32(sh0), wm_u = smem_u32(wm), dO_u = smem_u32(s_dO);
  const uint32_t rec_u = smem_u32(s_rec), bar_u = smem_u32(smem + lay.bar);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;           // mma fragment coordinates
  const int mi = lane >> 3, rr = lane & 7;           // ldmatrix: this lane addresses row rr of 8x8 matrix mi
  if (tid == 0) {
    mbar_init(bar_u, 1);
    mbar_init(bar_u + 8, 1);
    fence_barrier_init();
  }
  __syncthreads();

  // per-net kernel parameters picked with selects (a runtime-indexed array would live in local memory)
  auto net_h = [&](int net) { return reinterpret_cast<const __nv_bfloat16*>(net == 0 ? L.actor.h : net == 1 ? L.critic.h : L.linvel.h); };
  auto net_ldh = [&](int net) { return net == 0 ? L.actor.ld_h : net == 1 ? L.critic.ld_h : L.linvel.ld_h; };
  auto net_dz = [&](int net) { return reinterpret_cast<__nv_bfloat16*>(net == 0 ? L.actor.dZ : net == 1 ? L.critic.dZ : L.linvel.dZ); };
  auto net_lddz = [&](int net) { return net == 0 ? L.actor.ld_dz : net == 1 ? L.critic.ld_dz : L.linvel.ld_dz; };
  // Four consecutive weights per step, all of a thread's loads issued before the first is used (the element-wise loop
  // paid one L2 / DRAM round trip per iteration, 16 in a row: ~15 us of an 86 us launch).
  constexpr int STEPS = 2 * NO * HMAX / 4 / THREADS;
  const int h4 = H >> 2;
  float4 w4[STEPS];
  int dst[STEPS];
#pragma unroll
  for (int u = 0; u < STEPS; ++u) {
    const int i = tid + u * THREADS;                 // float4 index over [2 matrices][NO rows][H / 4]
    const int row = i / h4, k = (i - row * h4) * 4, mat = row / NO, n = row - mat * NO;
    const float* src = nullptr;
    if (mat == 0) { if (n < A) src = L.actor.W + n * H + k; }
    else if (mat == 1) { if (n == 0) src = L.critic.W + k; else if (n < 4) src = L.linvel.W + (n - 1) * H + k; }
    w4[u] = src ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
    dst[u] = mat < 2 ? (mat * 3 * NO + n) * HP + k : -1;
  }
  float b_early = 0.f, sd_early = 1.f;               // biases and std: loaded here, stored behind the first tile's issue
  if (tid < 3 * NO) {
    const int net = tid / NO, n = tid - net * NO;
    if (net == 0) { if (n < A) b_early = __ldg(L.actor.b + n); }
    else if (net == 1) { if (n == 0) b_early = __ldg(L.critic.b); }
    else if (n >= 1 && n < 4) b_early = __ldg(L.linvel.b + n - 1);
  }
  if (tid < A) sd_early = __ldg(L.std + tid);
  const int r = tid / SUB, sub = tid % SUB;
  const float inv_rows = L.inv_world / (float)L.rows;      // 1 / (rows over all ranks)
  const int ntiles = (L.rows + ROWS - 1) / ROWS;
  // Row of the old-policy record that thread (r, sub) reads for tile t (minibatch permutation), 0 past the end.
  auto record_row = [&](int t) -> long long {
    const int row = t * ROWS + r;
    if (t >= ntiles || row >= L.rows) return 0;
    if (!L.idx) return (long long)row;
    long long v;      // volatile asm: the load keeps its place in program order, a tile ahead of its use
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(L.idx + row));
    return v;
  };
  // Everything tile t needs, global -> shared buffer b: the three activation tiles (one TMA bulk copy per row, rows past
  // the end zero-filled) and the rows' old-policy records (cp.async).
  auto issue_loads = [&](int t, int b, long long src) {
    const int row0 = t * ROWS;
    const int valid = min(ROWS, L.rows - row0);
    const uint32_t sh_b = sh0_u + (uint32_t)(b * lay.sh_buf);
    if (tid == 0) mbar_arrive_expect_tx(bar_u + b * 8, (uint32_t)(3 * valid * H * 2));
    if (tid < 3 * ROWS) {                        // one activation row per thread (a warp per net): a bulk copy of H * 2 bytes
      const int net = tid / ROWS, row = tid - net * ROWS;
      const uint32_t d = sh_b + (uint32_t)((net * ROWS + row) * HP) * 2;
      if (row < valid) {
        bulk_g2s(d, net_h(net) + (long long)(row0 + row) * net_ldh(net), (uint32_t)H * 2, bar_u + b * 8);
      } else {                                   // rows past the end are zero (their gradients are, too: no NaN from stale bits)
        for (int c = 0; c < H * 2; c += 16) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d + c), "r"(0) : "memory");
      }
    }
    // the record: 3 * A / 4 16-byte pieces of the action-wise arrays (A % 4 == 0 on this path) and 7 scalars per row,
    // spread over the row's SUB lanes: 3 copy instructions per thread instead of 13 four-byte ones
    const bool rok = r < valid;
    const uint32_t rec_row = rec_u + (uint32_t)((b * ROWS + r) * RS) * 4;
    const int a4 = A >> 2;
#pragma unroll
    for (int c0 = 0; c0 < 3 * NO / 4; c0 += SUB) {
      const int c = c0 + sub, arr = c / a4, part = c - arr * a4;
      if (arr < 3) {
        const float* g = (arr == 0 ? L.actions : arr == 1 ? L.old_mu : L.old_sigma) + src * A + part * 4;
        cp_async16(rec_row + (uint32_t)(arr * NO + part * 4) * 4, g, rok);
      }
    }
    if (sub < 7) {
      const float* g = sub == 0 ? L.old_log_prob + src : sub == 1 ? L.advantages + src : sub == 2 ? L.returns + src
                     : sub == 3 ? L.old_values + src : L.critic_obs + src * L.ld_cobs + L.lv_col + (sub - 4);
      cp_async4(rec_row + (uint32_t)(RS_SCALARS + sub) * 4, g, rok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- software pipeline over the block's tiles: tile t + 1 is in flight (cp.async, second buffer) while tile t is
  // computed on, and the permutation index of tile t + 2 sits in a register.  Unpipelined, every block spent 7.5 of its
  // ~19 us per 64 rows waiting for loads, all blocks at the same time (device timeline, tools/bench_loss.py).
  issue_loads(blockIdx.x, 0, record_row(blockIdx.x));
  long long src_next = record_row(blockIdx.x + gridDim.x);
  // ---- once per block: weights as three bf16 terms, biases, per-action constants, zeroed gradient tiles
#pragma unroll
  for (int u = 0; u < STEPS; ++u) {
    if (dst[u] < 0) continue;
    const float w[4] = {w4[u].x, w4[u].y, w4[u].z, w4[u].w};
    __align__(8) __nv_bfloat16 t[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      t[0][e] = __float2bfloat16_rn(w[e]);
      const float r1 = w[e] - __bfloat162float(t[0][e]);
      t[1][e] = __float2bfloat16_rn(r1);
      t[2][e] = __float2bfloat16_rn(r1 - __bfloat162float(t[1][e]));
    }
#pragma unroll
    for (int sp = 0; sp < 3; ++sp)
      *reinterpret_cast<uint2*>(wm + (size_t)sp * NO * HP + dst[u]) = *reinterpret_cast<const uint2*>(t[sp]);
  }
  for (int i = tid; i < 3 * 2 * ROWS * DP / 2; i += THREADS) reinterpret_cast<uint32_t*>(s_dO)[i] = 0u;
  if (tid < 3 * NO) s_bias[tid] = b_early;
  if (tid < NO) {
    const float sd = sd_early;
    s_c[tid] = sd;
    s_c[NO + tid] = 1.f / (2.f * sd * sd);
    s_c[2 * NO + tid] = logf(sd);
    s_c[3 * NO + tid] = 1.f / sd;
  }
  __syncthreads();
  if (tid == 0) trace(0x001);

  // ---- accumulators that live for the whole launch
  float gWf[MAXP][2][4];      // [pair][n8 tile][fragment]: rows = outputs g8 / g8+8, cols = 2 t4, 2 t4 + 1
  float gbh[MAXP][4];         // column sums of dZ: cols {2 t4, 2 t4 + 1} of the two n8 tiles
#pragma unroll
  for (int j = 0; j < MAXP; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) gWf[j][0][i] = gWf[j][1][i] = gbh[j][i] = 0.f;
  }
  float st_v = 0.f, st_s = 0.f, st_l = 0.f, st_k = 0.f;
  float gb_a[AJ], gs_a[AJ], gb_c = 0.f, gb_l[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < AJ; ++j) gb_a[j] = gs_a[j] = 0.f;

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int row0 = tile * ROWS;
    const int valid = min(ROWS, L.rows - row0);
    const bool row_ok = r < valid;
    const int round = tile / (int)gridDim.x;
    if (tid == 0) trace(0x002, round);
    if (tile + (int)gridDim.x < ntiles) issue_loads(tile + gridDim.x, buf ^ 1, src_next);
    else asm volatile("cp.async.commit_group;" ::: "memory");     // (empty group: keeps the wait below uniform)
    src_next = record_row(tile + 2 * gridDim.x);
    if (tid == 0) trace(0x008, round);
    asm volatile("cp.async.wait_group 1;" ::: "memory");          // this tile's record has landed; the next one may be in flight
    mbar_wait(bar_u + buf * 8, (uint32_t)(it >> 1) & 1u);         // ... and its activation rows (k-th use of the buffer: parity k & 1)
    __syncthreads();
    if (tid == 0) trace(0x003, round);
    const __nv_bfloat16* sh = sh0 + (size_t)buf * (lay.sh_buf / 2);
    const uint32_t sh_u = sh0_u + (uint32_t)(buf * lay.sh_buf);
    const float* rec = s_rec + (size_t)(buf * ROWS + r) * RS;           // this thread's row
    const float* rowrec = rec + RS_SCALARS;

    // -- forward: out[net][ROWS][16] = h W^T.  Jobs = (net, 16-row group, n8 tile); the second n8 tile exists only for the
    //    actor (outputs 8..A-1): 8 jobs, one per warp.
    for (int job = warp; job < 4 * (ROWS / 16); job += WARPS) {
      const int q = job / (ROWS / 16), rg = job - q * (ROWS / 16);
      const int net = q < 2 ? 0 : q - 1, nt = q == 1 ? 1 : 0;
      // A: rows rg*16.., matrices {rows 0-7 | 8-15} x {k 0-7 | 8-15};  B: W[n][k], matrices {k 0-7 | 8-15} of two 16-wide k steps
      const uint32_t a_addr = sh_u + (uint32_t)((net * ROWS + rg * 16 + (mi & 1) * 8 + rr) * HP + (mi >> 1) * 8) * 2;
      const uint32_t b_addr = wm_u + (uint32_t)(((net == 0 ? 0 : 3) * NO + nt * 8 + rr) * HP + mi * 8) * 2;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k0 = 0; k0 < H; k0 += 32) {            // H % 32 == 0 on this path
        uint32_t a0[4], a1[4];
        ldsm_x4(a0, a_addr + k0 * 2);
        ldsm_x4(a1, a_addr + (k0 + 16) * 2);
#pragma unroll
        for (int sp = 2; sp >= 0; --sp) {               // smallest term first
          uint32_t b[4];                                // {b0, b1} of k step k0 and of k step k0 + 16
          ldsm_x4(b, b_addr + (uint32_t)(sp * NO * HP + k0) * 2);
          mma16816(c, a0, b[0], b[1]);
          mma16816(c, a1, b[2], b[3]);
        }
      }
      float* o = s_out + (size_t)(net * ROWS + rg * 16) * NO + nt * 8 + 2 * t4;
      *reinterpret_cast<float2*>(o + g8 * NO) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(o + (g8 + 8) * NO) = make_float2(c[2], c[3]);
    }
    __syncthreads();
    if (tid == 0) trace(0x004, round);

    // -- losses and d loss / d outputs (same math as ppo_head_loss_kernel; SUB lanes per row split the actions)
    {
      float logp = 0.f, kl = 0.f;
      float m[AJ];
#pragma unroll
      for (int j = 0; j < AJ; ++j) {
        const int a = sub + SUB * j;
        m[j] = 0.f;
        if (a < A && row_ok) {
          m[j] = s_out[r * NO + a] + s_bias[a];
          const float sd = s_c[a], i2v = s_c[NO + a], lsd = s_c[2 * NO + a];
          const float d = rec[a] - m[j];
          logp += -(d * d) * i2v - lsd - LOG_SQRT_2PI;
          const float osd = rec[2 * NO + a];
          const float dm = rec[NO + a] - m[j];
          kl += logf(sd * __frcp_rn(osd) + 1.e-5f) + (osd * osd + dm * dm) * i2v - 0.5f;
        }
      }
#pragma unroll
      for (int o = SUB >> 1; o > 0; o >>= 1) {
        logp += __shfl_xor_sync(0xffffffffu, logp, o);
        kl += __shfl_xor_sync(0xffffffffu, kl, o);
      }
      float g_logp = 0.f;
      __nv_bfloat16* dOa = s_dO;                                 // actor   [2][ROWS][DP]
      __nv_bfloat16* dOc = s_dO + (size_t)2 * ROWS * DP;         // critic
      __nv_bfloat16* dOl = s_dO + (size_t)4 * ROWS * DP;         // lin-vel
      auto put = [&](__nv_bfloat16* base, int col, float g) {    // gradient as hi + lo bf16 terms
        const __nv_bfloat16 hi = __float2bfloat16_rn(g);
        base[r * DP + col] = hi;
        base[ROWS * DP + r * DP + col] = __float2bfloat16_rn(g - __bfloat162float(hi));
      };
      if (row_ok) {
        const float r_adv = rowrec[1];
        const float ratio = expf(logp - rowrec[0]);
        const float s1 = -r_adv * ratio;
        const float s2 = -r_adv * fminf(fmaxf(ratio, 1.f - L.clip_param), 1.f + L.clip_param);
        g_logp = (s1 >= s2 ? s1 : 0.f) * inv_rows;               // d surr / d logp  (ties: clamp passes gradient)
        if (sub == 0) {
          st_s += fmaxf(s1, s2); st_k += kl;
          const float R = rowrec[2], v = s_out[ROWS * NO + r * NO] + s_bias[NO];
          float vl, g_v;
          if (L.use_clipped_value_loss) {
            const float vo = rowrec[3];
            const float dv = v - vo;
            const float vc = vo + fminf(fmaxf(dv, -L.clip_param), L.clip_param);
            const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
            const float pass = (dv >= -L.clip_param && dv <= L.clip_param) ? 1.f : 0.f;
            vl = fmaxf(l1, l2);
            g_v = l1 > l2 ? 2.f * (v - R) : (l1 < l2 ? 2.f * (vc - R) * pass : (v - R) + (vc - R) * pass);
          } else {
            vl = (R - v) * (R - v);
            g_v = 2.f * (v - R);
          }
          st_v += vl;
          const float gv = L.value_loss_coef * g_v * inv_rows;
          put(dOc, 0, gv);
          gb_c += gv;
          float lsq = 0.f;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float e = s_out[2 * ROWS * NO + r * NO + 1 + j] + s_bias[2 * NO + 1 + j] - rowrec[4 + j];
            lsq += e * e;
            const float gl = L.lin_vel_coef * 2.f * e * inv_rows * (1.f / 3.f);
            put(dOl, 1 + j, gl);
            gb_l[j] += gl;
          }
          st_l += lsq;
        }
      } else if (sub == 0) {
        put(dOc, 0, 0.f);
#pragma unroll
        for (int j = 0; j < 3; ++j) put(dOl, 1 + j, 0.f);
      }
#pragma unroll
      for (int j = 0; j < AJ; ++j) {
        const int a = sub + SUB * j;
        if (a < A) {
          float gm = 0.f;
          if (row_ok) {
            const float isd = s_c[3 * NO + a], ivar = 2.f * s_c[NO + a];
            const float d = rec[a] - m[j];
            gm = g_logp * d * ivar;
            gs_a[j] += g_logp * (d * d * ivar * isd - isd) - L.entropy_coef * inv_rows * isd;
          }
          put(dOa, a, gm);
          gb_a[j] += gm;
        }
      }
    }
    __syncthreads();
    if (tid == 0) trace(0x005, round);

    // -- backward: each warp owns up to MAXP (net, 16-column tile) pairs for the whole launch
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      const int p = warp + WARPS * j;
      if (p < P) {
        const int net = p / NT, n0 = (p - net * NT) * 16;
        const uint32_t h_u = sh_u + (uint32_t)(net * ROWS * HP) * 2;
        const uint32_t d_u = dO_u + (uint32_t)(net * 2 * ROWS * DP) * 2;
        const uint32_t w_u = wm_u + (uint32_t)((net == 0 ? 0 : 3) * NO * HP) * 2;
        // gW[16 outputs x 16 cols] += dOut^T[16 x ROWS] h[ROWS x 16 cols]:
        //   A(m = output, k = row) is dOut[row][output]  -> transposed 8x8 loads, matrices {k 0-7 | 8-15} x {m 0-7 | 8-15}
        //   B(k = row, n = col)    is h[row][col]        -> transposed 8x8 loads
        for (int k0 = 0; k0 < ROWS; k0 += 16) {
          uint32_t b[4], ah[4], al[4];
          ldsm_x4_t(b, h_u + (uint32_t)((k0 + (mi & 1) * 8 + rr) * HP + n0 + (mi >> 1) * 8) * 2);
          const uint32_t a_off = (uint32_t)((k0 + (mi >> 1) * 8 + rr) * DP + (mi & 1) * 8) * 2;
          ldsm_x4_t(ah, d_u + a_off);
          ldsm_x4_t(al, d_u + (uint32_t)(ROWS * DP) * 2 + a_off);
          mma16816(gWf[j][0], al, b[0], b[1]);
          mma16816(gWf[j][1], al, b[2], b[3]);
          mma16816(gWf[j][0], ah, b[0], b[1]);
          mma16816(gWf[j][1], ah, b[2], b[3]);
        }
        // dH[16 rows x 16 cols] = dOut[16 x 16 outputs] W[16 outputs x 16 cols] per 16-row group (B: transposed loads of
        // W[output][col]), then the ELU derivative, the dZ store and the bias-gradient column sums -- all in fragments
        uint32_t bh[4], bm[4];
        const uint32_t w_off = (uint32_t)(((mi & 1) * 8 + rr) * HP + n0 + (mi >> 1) * 8) * 2;
        ldsm_x4_t(bh, w_u + w_off);
        ldsm_x4_t(bm, w_u + (uint32_t)(NO * HP) * 2 + w_off);
        const __nv_bfloat16* h_lane = sh + (size_t)(net * ROWS + g8) * HP + n0 + 2 * t4;    // this lane's first (row, col)
        const long long ldz8 = 8ll * net_lddz(net);
        __nv_bfloat16* dz_lane = net_dz(net) + (long long)(row0 + g8) * net_lddz(net) + n0 + 2 * t4;   // stepped 8 rows at a time
        const int rows_left = valid - g8;
#pragma unroll
        for (int rg = 0; rg < ROWS / 16; ++rg) {
          uint32_t ah[4], al[4];
          const uint32_t a_off = (uint32_t)((rg * 16 + (mi & 1) * 8 + rr) * DP + (mi >> 1) * 8) * 2;
          ldsm_x4(ah, d_u + a_off);
          ldsm_x4(al, d_u + (uint32_t)(ROWS * DP) * 2 + a_off);
          float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
          mma16816(c[0], ah, bm[0], bm[1]);
          mma16816(c[1], ah, bm[2], bm[3]);
          mma16816(c[0], al, bh[0], bh[1]);
          mma16816(c[1], al, bh[2], bh[3]);
          mma16816(c[0], ah, bh[0], bh[1]);
          mma16816(c[1], ah, bh[2], bh[3]);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf, dz_lane += ldz8) {   // rows g8 and g8 + 8
            const int off = rg * 16 + hf * 8;                 // row offset from this lane's first row (g8)
            const bool ok = off < rows_left;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const float2 h2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(h_lane + off * HP + nt * 8));
              const float z0 = c[nt][2 * hf] * (h2.x > 0.f ? 1.f : h2.x + 1.f);
              const float z1 = c[nt][2 * hf + 1] * (h2.y > 0.f ? 1.f : h2.y + 1.f);
              gbh[j][2 * nt] += z0;
              gbh[j][2 * nt + 1] += z1;
              if (ok) *reinterpret_cast<uint32_t*>(dz_lane + nt * 8) = pack_bf16x2(z0, z1);
            }
          }
        }
      }
    }
    __syncthreads();       // the next tile overwrites sh / s_out / s_dO
    if (tid == 0) trace(0x006, round);
  }
  if (tid == 0) pdl_launch_dependents();     // persistent grid, every block resident: the successor's launch overlaps the flush

  // ---- once per block: flush everything that was accumulated in registers
#pragma unroll
  for (int j = 0; j < MAXP; ++j) {
    const int p = warp + WARPS * j;
    if (p < P) {
      const int net = p / NT, n0 = (p - net * NT) * 16;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int a = g8 + (i >> 1) * 8, col = n0 + nt * 8 + 2 * t4 + (i & 1);
          const float g = gWf[j][nt][i];
          if (net == 0) { if (a < A) atomicAdd(L.actor.gW + a * H + col, g); }
          else if (net == 1) { if (a == 0) atomicAdd(L.critic.gW + col, g); }
          else if (a >= 1 && a < 4) atomicAdd(L.linvel.gW + (a - 1) * H + col, g);
        }
      }
      float* gbh_g = net == 0 ? L.actor.gb_hidden : net == 1 ? L.critic.gb_hidden : L.linvel.gb_hidden;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = gbh[j][i];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (g8 == 0) atomicAdd(gbh_g + n0 + (i >> 1) * 8 + 2 * t4 + (i & 1), v);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < AJ; ++j) {
    float g = gb_a[j], gs = gs_a[j];
#pragma unroll
    for (int o = SUB; o < 32; o <<= 1) {
      g += __shfl_xor_sync(0xffffffffu, g, o);
      gs += __shfl_xor_sync(0xffffffffu, gs, o);
    }
    const int a = lane + SUB * j;
    if (lane < SUB && a < A) {
      atomicAdd(L.actor.gb + a, g);
      atomicAdd(L.g_std + a, gs);
    }
  }
  gb_c = warp_sum(gb_c);
#pragma unroll
  for (int j = 0; j < 3; ++j) gb_l[j] = warp_sum(gb_l[j]);
  st_v = warp_sum(st_v); st_s = warp_sum(st_s); st_l = warp_sum(st_l); st_k = warp_sum(st_k);
  if (lane == 0) {
    atomicAdd(L.critic.gb, gb_c);
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(L.linvel.gb + j, gb_l[j]);
    atomicAdd(L.stats + 0, st_v);
    atomicAdd(L.stats + 1, st_s);
    atomicAdd(L.stats + 2, st_l);
    atomicAdd(L.stats + 3, st_k);
  }
  // The entropy of a diagonal Gaussian does not depend on the row: rows * sum_a (1/2 + log sqrt(2 pi) + log sigma_a), added
  // once.  (Summed row by row in fp32 it drifts: every block adds the SAME value, so the rounding of the adds is a bias.)
  if (blockIdx.x == 0 && tid == 0) {
    double e = 0.0;
    for (int a = 0; a < A; ++a) e += (double)(HALF_LOG_2PI_PLUS_HALF + s_c[2 * NO + a]);
    atomicAdd(L.stats + 4, (float)(e * (double)L.rows));
  }
  if (tid == 0) trace(0x007);
}
HG_TRACE_SETTER(trace_set_heads)

static bool loss_mma_eligible(const HgLossArgs& L) {
  const int H = L.actor.H;
  const uintptr_t rec_align = reinterpret_cast<uintptr_t>(L.actions) | reinterpret_cast<uintptr_t>(L.old_mu) | reinterpret_cast<uintptr_t>(L.old_sigma);
  const uintptr_t w_align = reinterpret_cast<uintptr_t>(L.actor.W) | reinterpret_cast<uintptr_t>(L.critic.W) | reinterpret_cast<uintptr_t>(L.linvel.W);
  const uintptr_t h_align = reinterpret_cast<uintptr_t>(L.actor.h) | reinterpret_cast<uintptr_t>(L.critic.h) | reinterpret_cast<uintptr_t>(L.linvel.h) |
                            (uintptr_t)((L.actor.ld_h | L.critic.ld_h | L.linvel.ld_h) & 7) * 2;
  return H == L.critic.H && H == L.linvel.H && H % 32 == 0 && H >= 32 && H <= mh::HMAX && L.A <= mh::NO && L.A % 4 == 0 && ((w_align | rec_align | h_align) & 15) == 0;
}

static int launch_loss_mma(const HgLossArgs& L, cudaStream_t s) {
  const size_t smem = mh::Layout(L.actor.H).total;
  auto kernel = ppo_head_loss_mma_kernel;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    HG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int ntiles = (L.rows + mh::ROWS - 1) / mh::ROWS;
  const int cap = 2 * num_sms();
  HG_CUDA(launch_kernel(kernel, dim3(ntiles < cap ? ntiles : cap), dim3(mh::THREADS), smem, s, 1, L));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

static size_t loss_smem_bytes(const HgLossArgs& L, int AT) {
  auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
  size_t t = 0;
  const HgHeadNet* nets[3] = {&L.actor, &L.critic, &L.linvel};
  for (int i = 0; i < 3; ++i) {
    t += al((size_t)LH_ROWS * (nets[i]->H + 8) * 2) + al((size_t)nets[i]->n_out * nets[i]->H * 4) +
         al((size_t)nets[i]->n_out * 4) + al((size_t)LH_ROWS * nets[i]->n_out * 4);
  }
  t += al(MAX_A * 4) + al(3 * MAX_A * 4) + al((size_t)LH_ROWS * (3 * L.A + 8) * 4) + al((size_t)LH_ROWS * L.A * 4) +
       al((size_t)(LH_THREADS / 32) * 8 * 4) + al((size_t)LH_ROWS * 8) + al((size_t)(LH_GROUPS - 1) * 128 * (AT + 1) * 4 + 16);
  return t;
}

template <int AT, bool EXACT>
static int launch_loss(const HgLossArgs& L, cudaStream_t s) {
  const size_t smem = loss_smem_bytes(L, AT);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    HG_CUDA(cudaFuncSetAttribute(ppo_head_loss_kernel<AT, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  HG_CUDA(launch_kernel(ppo_head_loss_kernel<AT, EXACT>, dim3((L.rows + LH_ROWS - 1) / LH_ROWS), dim3(LH_THREADS), smem, s, 1, L));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

template <int AT>
static int launch_rollout(const void* h_actor, int64_t ld_ha, int H_a, const float* W_a, const float* b_a, const float* std,
                          const void* h_critic, int64_t ld_hc, int H_c, const float* W_c, const float* b_c, const float* eps,
                          uint64_t seed, uint64_t counter, int n, int A, float* actions, float* mu, float* sigma,
                          float* log_prob, float* values, cudaStream_t s) {
  const size_t smem = (size_t)(A * H_a + H_c + 2 * MAX_A + 4) * 4 + (size_t)RH_ROWS * ((h_actor ? H_a + 8 : 0) + H_c + 8) * 2 + 32;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    HG_CUDA(cudaFuncSetAttribute(policy_head_rollout_kernel<AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  HG_CUDA(launch_kernel(policy_head_rollout_kernel<AT>, dim3((n + RH_ROWS - 1) / RH_ROWS), dim3(RH_ROWS * RH_SUB), smem, s, 1, reinterpret_cast<const __nv_bfloat16*>(h_actor), ld_ha, H_a, W_a, b_a, std, reinterpret_cast<const __nv_bfloat16*>(h_critic), ld_hc, H_c, W_c, b_c, eps, seed, counter, n, A, actions, mu, sigma, log_prob, values));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

}  // namespace hg

namespace hg {
// Generic `ActorCritic.act` / `get_actions_log_prob` (actor_critic.py:127-138) for callers outside the fused rollout path:
// a = mu + std * eps (eps given, or Philox4x32-10 keyed by (seed, counter, row, action)) and sum_a log N(a | mu, std).
// One thread per row; the same fp32 expressions as the fused heads.
__global__ void gaussian_sample_kernel(const float* __restrict__ mu, const float* __restrict__ std, const float* __restrict__ eps,
                                       uint64_t seed, uint64_t counter, int n, int A, float* __restrict__ actions,
                                       float* __restrict__ log_prob) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  float lp = 0.f;
  for (int a = 0; a < A; ++a) {
    const float m = mu[(long long)row * A + a], sd = __ldg(std + a);
    const float e = eps ? eps[(long long)row * A + a] : philox_normal(seed, counter, (uint32_t)row, (uint32_t)a);
    const float act = m + sd * e;
    const float d = act - m;
    lp += -(d * d) / (2.f * sd * sd) - logf(sd) - LOG_SQRT_2PI;
    actions[(long long)row * A + a] = act;
  }
  if (log_prob) log_prob[row] = lp;
}

__global__ void gaussian_log_prob_kernel(const float* __restrict__ actions, const float* __restrict__ mu, const float* __restrict__ std,
                                         int n, int A, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  float lp = 0.f;
  for (int a = 0; a < A; ++a) {
    const float sd = __ldg(std + a);
    const float d = actions[(long long)row * A + a] - mu[(long long)row * A + a];
    lp += -(d * d) / (2.f * sd * sd) - logf(sd) - LOG_SQRT_2PI;
  }
  out[row] = lp;
}
}  // namespace hg

extern "C" int hg_gaussian_sample(const float* mu, const float* std, const float* eps, uint64_t seed, uint64_t counter, int n, int A,
                                  float* actions, float* log_prob, void* stream) {
  using namespace hg;
  HG_CHECK(mu && std && actions && n > 0 && A >= 1, "hg_gaussian_sample: bad arguments");
  HG_CUDA(launch_kernel(gaussian_sample_kernel, dim3((n + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 1, mu, std, eps,
                        seed, counter, n, A, actions, log_prob));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_gaussian_log_prob(const float* actions, const float* mu, const float* std, int n, int A, float* out, void* stream) {
  using namespace hg;
  HG_CHECK(actions && mu && std && out && n > 0 && A >= 1, "hg_gaussian_log_prob: bad arguments");
  HG_CUDA(launch_kernel(gaussian_log_prob_kernel, dim3((n + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 1, actions, mu,
                        std, n, A, out));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_policy_head_rollout(const void* h_actor, int64_t ld_ha, int H_a, const float* W_a, const float* b_a,
                                      const float* std, const void* h_critic, int64_t ld_hc, int H_c, const float* W_c,
                                      const float* b_c, const float* eps, uint64_t seed, uint64_t counter, int n, int A,
                                      float* actions, float* mu, float* sigma, float* log_prob, float* values, void* stream) {
  using namespace hg;
  HG_CHECK(h_critic && W_c && b_c && values && n > 0, "hg_policy_head_rollout: bad critic arguments");
  HG_CHECK(H_c % 8 == 0 && H_c <= 512 && ld_hc % 8 == 0, "critic hidden width must be a multiple of 8, <= 512");
  if (h_actor) {
    HG_CHECK(W_a && b_a && std && actions && mu && sigma && log_prob, "hg_policy_head_rollout: null actor pointer");
    HG_CHECK(A >= 1 && A <= MAX_A, "num_actions must be in [1,%d]", MAX_A);
    HG_CHECK(H_a % 8 == 0 && H_a <= 512 && ld_ha % 8 == 0, "actor hidden width must be a multiple of 8, <= 512");
  } else {
    H_a = 0;
    A = 0;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (A <= 12)
    return launch_rollout<12>(h_actor, ld_ha, H_a, W_a, b_a, std, h_critic, ld_hc, H_c, W_c, b_c, eps, seed, counter, n, A,
                              actions, mu, sigma, log_prob, values, s);
  if (A <= 20)
    return launch_rollout<20>(h_actor, ld_ha, H_a, W_a, b_a, std, h_critic, ld_hc, H_c, W_c, b_c, eps, seed, counter, n, A,
                              actions, mu, sigma, log_prob, values, s);
  return launch_rollout<32>(h_actor, ld_ha, H_a, W_a, b_a, std, h_critic, ld_hc, H_c, W_c, b_c, eps, seed, counter, n, A,
                            actions, mu, sigma, log_prob, values, s);
}

extern "C" int hg_value_head(const void* h_critic, int64_t ld_hc, int H_c, const float* W_c, const float* b_c, int n,
                             float* values, void* stream) {
  return hg_policy_head_rollout(nullptr, 0, 0, nullptr, nullptr, nullptr, h_critic, ld_hc, H_c, W_c, b_c, nullptr, 0, 0,
                                n, 0, nullptr, nullptr, nullptr, nullptr, values, stream);
}

extern "C" int hg_head_forward(const void* h, int64_t ld_h, int H, const float* W, const float* b, int n_out, int n,
                               float* out, void* stream) {
  using namespace hg;
  HG_CHECK(h && W && b && out && n > 0, "hg_head_forward: bad arguments");
  HG_CHECK(n_out >= 1 && n_out <= MAX_A, "hg_head_forward: n_out must be in [1,%d]", MAX_A);
  HG_CHECK(H % 8 == 0 && H >= 8 && H <= 512 && ld_h % 8 == 0, "hg_head_forward: hidden width must be a multiple of 8, <= 512");
  const size_t smem = (size_t)(n_out * H + MAX_A) * 4 + (size_t)RH_ROWS * (H + 8) * 2 + 32;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    HG_CUDA(cudaFuncSetAttribute(head_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  HG_CUDA(launch_kernel(head_forward_kernel, dim3((n + RH_ROWS - 1) / RH_ROWS), dim3(RH_ROWS * RH_SUB), smem, reinterpret_cast<cudaStream_t>(stream), 1, reinterpret_cast<const __nv_bfloat16*>(h), ld_h, H, W, b, n_out, n, out));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_ppo_head_loss(const HgLossArgs* L, void* stream) {
  using namespace hg;
  HG_CHECK(L != nullptr && L->rows > 0, "hg_ppo_head_loss: bad arguments");
  HG_CHECK(L->A >= 1 && L->A <= MAX_A && L->actor.n_out == L->A, "num_actions must be in [1,%d] and match the actor head", MAX_A);
  HG_CHECK(L->critic.n_out == 1 && L->linvel.n_out == 3, "critic head must have 1 output, lin-vel head 3");
  const HgHeadNet* nets[3] = {&L->actor, &L->critic, &L->linvel};
  for (int i = 0; i < 3; ++i) {
    HG_CHECK(nets[i]->H % 8 == 0 && nets[i]->H >= 8 && nets[i]->H <= 256, "head %d: hidden width must be a multiple of 8 in [8,256]", i);
    HG_CHECK(nets[i]->h && nets[i]->W && nets[i]->b && nets[i]->gW && nets[i]->gb && nets[i]->dZ && nets[i]->gb_hidden, "head %d: null pointer", i);
    HG_CHECK(nets[i]->ld_h % 8 == 0, "head %d: ld_h alignment", i);
  }
  HG_CHECK(L->actions && L->old_mu && L->old_sigma && L->old_log_prob && L->advantages && L->returns && L->old_values &&
               L->critic_obs && L->std && L->g_std && L->stats, "hg_ppo_head_loss: null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static const bool simt_only = getenv("HG_LOSS_SIMT") != nullptr;     // A/B switch for the tests
  if (!simt_only && loss_mma_eligible(*L)) return launch_loss_mma(*L, s);
  if (L->A == 12) return launch_loss<12, true>(*L, s);     // XBot-L: 12 actions
  if (L->A == 20) return launch_loss<20, true>(*L, s);
  if (L->A <= 12) return launch_loss<12, false>(*L, s);
  if (L->A <= 20) return launch_loss<20, false>(*L, s);    // 18-DOF fork
  return launch_loss<32, false>(*L, s);
}
