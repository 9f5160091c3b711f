// Mirror-symmetry loss of PPO.update (reference ppo.py:93-105,197-202; off in the shipped config):
//   mirror_obs   = obs_batch @ obs_perm_mat              -> permute_cols_kernel (signed column gather of the bf16 rows)
//   mirror_act   = actor(mirror_obs)                      -> the grouped tcgen05 GEMMs + hg_head_forward (existing kernels)
//   sym_loss     = mean((mu - mirror_act @ act_perm_mat)^2)   -> sym_loss_kernel (+ d/d mu, d/d mirror_act)
//   backward of the actor's output layer for an extra d/d out    -> head_backward_kernel (dZ of the top hidden layer,
//                 output-layer weight / bias gradients, top hidden bias gradient); the hidden layers then go through
//                 the same grouped dW / dX GEMMs as the main path.
// A signed permutation matrix P (P[|p_i|][i] = sign(p_i)) is passed as src[i] = |p_i| and sign[i] in {-1, 0, +1}
// (0: np.sign(0) -- the reference zeroes that column).
#include "hg_common.cuh"

namespace hg {

// One warp per row: dst[r, j] = sign[j] * src[r, src_col[j]] (bf16: a sign flip is an XOR), zero padding up to ld_dst.
__global__ void permute_cols_kernel(const uint16_t* __restrict__ src, long long ld_src, int rows, int cols,
                                    const int* __restrict__ src_col, const float* __restrict__ sign, uint16_t* __restrict__ dst,
                                    long long ld_dst) {
  pdl_trigger();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const uint16_t* s = src + (long long)r * ld_src;
    uint16_t* d = dst + (long long)r * ld_dst;
    for (int j = lane; j < (int)ld_dst; j += 32) {
      uint16_t v = 0;
      if (j < cols) {
        const float sg = __ldg(sign + j);
        if (sg != 0.f) v = s[__ldg(src_col + j)] ^ (sg < 0.f ? (uint16_t)0x8000u : (uint16_t)0u);
      }
      d[j] = v;
    }
  }
}

// One thread per row.  m_a = sign_a * mirror[r, src_a]; diff = mu - m; loss += diff^2;
// dmu[r, a] = scale * diff; dmirror[r, src_a] = -scale * sign_a * diff (a permutation: every column written at most once).
__global__ void sym_loss_kernel(const float* __restrict__ mu, const float* __restrict__ mirror, const int* __restrict__ a_src,
                                const float* __restrict__ a_sign, int rows, int A, float scale, float* __restrict__ dmu,
                                float* __restrict__ dmirror, double* __restrict__ loss_sum) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (r < rows) {
    for (int a = 0; a < A; ++a) dmirror[(long long)r * A + a] = 0.f;
    for (int a = 0; a < A; ++a) {
      const float sg = __ldg(a_sign + a);
      const int k = __ldg(a_src + a);
      const float m = sg * mirror[(long long)r * A + k];
      const float d = mu[(long long)r * A + a] - m;
      acc += d * d;
      dmu[(long long)r * A + a] = scale * d;
      if (sg != 0.f) dmirror[(long long)r * A + k] = -scale * sg * d;
    }
  }
  acc = warp_sum(acc);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += (double)s[w];
    atomicAdd(loss_sum, t);
  }
}

// Backward of an fp32 output layer out = h W^T + b for a given d/d out, one thread per hidden unit k, HB_ROWS rows per block:
//   dH[r,k] = sum_a dOut[r,a] W[a,k];  dZ[r,k] (+)= dH[r,k] * elu'(h[r,k]);  gb_hidden[k] += sum_r of that;
//   gW[a,k] += sum_r dOut[r,a] h[r,k];  gb[a] += sum_r dOut[r,a]
constexpr int HB_ROWS = 64;
constexpr int HB_MAX_OUT = 16;
__global__ void head_backward_kernel(const __nv_bfloat16* __restrict__ h, long long ld_h, int H, const float* __restrict__ W,
                                     int n_out, const float* __restrict__ dOut, int rows, int accumulate,
                                     __nv_bfloat16* __restrict__ dZ, long long ld_dz, float* __restrict__ gW,
                                     float* __restrict__ gb, float* __restrict__ gb_hidden) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_d[HB_ROWS * HB_MAX_OUT];
  const int r0 = blockIdx.x * HB_ROWS;
  const int nr = min(HB_ROWS, rows - r0);
  for (int i = threadIdx.x; i < nr * n_out; i += blockDim.x) s_d[i] = dOut[(long long)r0 * n_out + i];
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float w[HB_MAX_OUT], gw[HB_MAX_OUT];
#pragma unroll
    for (int a = 0; a < HB_MAX_OUT; ++a) {
      w[a] = a < n_out ? __ldg(W + (long long)a * H + k) : 0.f;
      gw[a] = 0.f;
    }
    float gbh = 0.f;
    for (int r = 0; r < nr; ++r) {
      const float hv = __bfloat162float(h[(long long)(r0 + r) * ld_h + k]);
      float dh = 0.f;
#pragma unroll
      for (int a = 0; a < HB_MAX_OUT; ++a) {
        if (a < n_out) {
          const float d = s_d[r * n_out + a];
          dh += d * w[a];
          gw[a] += d * hv;
        }
      }
      const float dz = dh * (hv > 0.f ? 1.f : hv + 1.f);
      __nv_bfloat16* out = dZ + (long long)(r0 + r) * ld_dz + k;
      const float prev = accumulate ? __bfloat162float(*out) : 0.f;
      *out = __float2bfloat16_rn(prev + dz);
      gbh += dz;
    }
#pragma unroll
    for (int a = 0; a < HB_MAX_OUT; ++a)
      if (a < n_out) atomicAdd(gW + (long long)a * H + k, gw[a]);
    if (gb_hidden) atomicAdd(gb_hidden + k, gbh);
  }
  if (gb && threadIdx.x < n_out) {
    float t = 0.f;
    for (int r = 0; r < nr; ++r) t += s_d[r * n_out + threadIdx.x];
    atomicAdd(gb + threadIdx.x, t);
  }
}

}  // namespace hg

extern "C" int hg_permute_cols_bf16(const void* src, int64_t ld_src, int rows, int cols, const int32_t* src_col,
                                    const float* sign, void* dst, int64_t ld_dst, void* stream) {
  using namespace hg;
  HG_CHECK(src && dst && src_col && sign && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols,
           "hg_permute_cols_bf16: bad arguments");
  long long blocks = ((long long)rows + 7) / 8;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  HG_CUDA(launch_kernel(permute_cols_kernel, dim3((unsigned)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                        reinterpret_cast<const uint16_t*>(src), (long long)ld_src, rows, cols, (const int*)src_col, sign,
                        reinterpret_cast<uint16_t*>(dst), (long long)ld_dst));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_sym_loss(const float* mu, const float* mirror_act, const int32_t* act_src, const float* act_sign, int rows,
                           int num_actions, float scale, float* dmu, float* dmirror, double* loss_sum, void* stream) {
  using namespace hg;
  HG_CHECK(mu && mirror_act && act_src && act_sign && dmu && dmirror && loss_sum && rows > 0 && num_actions > 0,
           "hg_sym_loss: bad arguments");
  HG_CUDA(launch_kernel(sym_loss_kernel, dim3((rows + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 1, mu,
                        mirror_act, (const int*)act_src, act_sign, rows, num_actions, scale, dmu, dmirror, loss_sum));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_head_backward(const void* h, int64_t ld_h, int H, const float* W, int n_out, const float* d_out, int rows,
                                int accumulate, void* dZ, int64_t ld_dz, float* gW, float* gb, float* gb_hidden, void* stream) {
  using namespace hg;
  HG_CHECK(h && W && d_out && dZ && gW && rows > 0 && H > 0 && n_out >= 1 && n_out <= HB_MAX_OUT,
           "hg_head_backward: bad arguments (n_out <= %d)", HB_MAX_OUT);
  HG_CUDA(launch_kernel(head_backward_kernel, dim3((rows + HB_ROWS - 1) / HB_ROWS), dim3(128), 0,
                        reinterpret_cast<cudaStream_t>(stream), 1, reinterpret_cast<const __nv_bfloat16*>(h), (long long)ld_h, H, W,
                        n_out, d_out, rows, accumulate, reinterpret_cast<__nv_bfloat16*>(dZ), (long long)ld_dz, gW, gb, gb_hidden));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}
