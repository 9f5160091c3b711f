// GAE reverse-time scan + advantage normalisation.
// Replaces RolloutStorage.compute_returns (reference humanoid/algo/ppo/rollout_storage.py:122-143).
//
// Layout: all tensors are [T, N] with N contiguous, so one thread per env gives perfectly coalesced
// 128-byte warp accesses at every time step.  The recurrence is sequential in t by construction (and
// must stay so to be bit-exact with the reference's fp32 op order), so latency is hidden by staging a
// whole [CHUNK_T x block] slab of rewards/values/dones in shared memory with cp.async before each run of
// CHUNK_T dependent steps: with CHUNK_T = 64 the XBot horizon (60) is ONE global round trip per thread.  Blocks are a
// single warp until there are enough envs to give every SM several warps (4096 envs: 128 blocks instead of the 32
// 128-thread blocks of round 1, which left 116 of the 148 SMs idle).  HBM traffic: 9 B read + 8 B written per (t, env)
// in the scan, 4 B + 4 B in the normalisation pass.
#include "hg_common.cuh"

namespace hg {

constexpr int GAE_CHUNK = 64;

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int GAE_THREADS>
__global__ void __launch_bounds__(GAE_THREADS) gae_scan_kernel(const float* __restrict__ rewards,
                                                               const float* __restrict__ values,
                                                               const uint8_t* __restrict__ dones,
                                                               const float* __restrict__ last_values, float gamma,
                                                               float lam, int T, int N, float* __restrict__ returns,
                                                               float* __restrict__ adv_raw, double* scratch) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_r[GAE_CHUNK][GAE_THREADS];
  __shared__ float s_v[GAE_CHUNK][GAE_THREADS];
  __shared__ uint8_t s_d[GAE_CHUNK][GAE_THREADS];
  __shared__ double s_red[2][GAE_THREADS / 32];

  const int n = blockIdx.x * GAE_THREADS + threadIdx.x;
  const bool ok = n < N;
  float adv = 0.f;
  float next_v = ok ? last_values[n] : 0.f;
  double sum = 0.0, sumsq = 0.0;

  for (int t_hi = T; t_hi > 0; t_hi -= GAE_CHUNK) {
    const int t_lo = max(0, t_hi - GAE_CHUNK);
    if (ok) {
      for (int t = t_lo; t < t_hi; ++t) {
        cp_async4(&s_r[t - t_lo][threadIdx.x], rewards + (size_t)t * N + n);
        cp_async4(&s_v[t - t_lo][threadIdx.x], values + (size_t)t * N + n);
        s_d[t - t_lo][threadIdx.x] = dones[(size_t)t * N + n];
      }
    }
    cp_async_commit_wait_all();   // each thread only reads back what it staged itself: no block barrier needed
    if (ok) {
      for (int t = t_hi - 1; t >= t_lo; --t) {
        const float r = s_r[t - t_lo][threadIdx.x];
        const float v = s_v[t - t_lo][threadIdx.x];
        // exact fp32 op order of the reference, no FMA contraction:
        //   not_term = 1 - done;  delta = r + not_term*gamma*next_v - v;  adv = delta + not_term*gamma*lam*adv
        const float not_term = __fsub_rn(1.0f, (float)s_d[t - t_lo][threadIdx.x]);
        const float ng = __fmul_rn(not_term, gamma);
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(ng, next_v)), v);
        adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(ng, lam), adv));
        const float ret = __fadd_rn(adv, v);
        returns[(size_t)t * N + n] = ret;
        const float a = __fsub_rn(ret, v);          // advantages = returns - values (:142)
        adv_raw[(size_t)t * N + n] = a;
        sum += (double)a;
        sumsq += (double)a * (double)a;
        next_v = v;
      }
    }
  }
  // block reduction of the two moments, one double atomic pair per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = sum;
    s_red[1][threadIdx.x >> 5] = sumsq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < GAE_THREADS / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
    atomicAdd(&scratch[0], a);
    atomicAdd(&scratch[1], b);
  }
}

__global__ void gae_normalise_kernel(float* __restrict__ adv, long long count, const double* __restrict__ scratch) {
  pdl_trigger();
  pdl_wait();
  const double n = (double)count;
  const double mean_d = scratch[0] / n;
  double var = (scratch[1] - scratch[0] * mean_d) / (n - 1.0);   // unbiased (torch.std default)
  var = var > 0.0 ? var : 0.0;
  const float mean = (float)mean_d;
  const float denom = __fadd_rn((float)sqrt(var), 1e-8f);
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += stride) {
    if (i + 4 <= count) {
      float4 a = *reinterpret_cast<float4*>(adv + i);
      a.x = __fdiv_rn(__fsub_rn(a.x, mean), denom);
      a.y = __fdiv_rn(__fsub_rn(a.y, mean), denom);
      a.z = __fdiv_rn(__fsub_rn(a.z, mean), denom);
      a.w = __fdiv_rn(__fsub_rn(a.w, mean), denom);
      *reinterpret_cast<float4*>(adv + i) = a;
    } else {
      for (long long j = i; j < count; ++j) adv[j] = __fdiv_rn(__fsub_rn(adv[j], mean), denom);
    }
  }
}

}  // namespace hg

extern "C" int hg_gae_returns(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                              float gamma, float lam, int T, int N, float* returns, float* advantages, double* scratch,
                              void* stream) {
  using namespace hg;
  HG_CHECK(rewards && values && dones && last_values && returns && advantages && scratch, "hg_gae_returns: null pointer");
  HG_CHECK(T > 0 && N > 0, "hg_gae_returns: empty rollout");
  HG_CHECK((reinterpret_cast<uintptr_t>(advantages) & 15) == 0, "advantages must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  HG_CUDA(cudaMemsetAsync(scratch, 0, 4 * sizeof(double), s));
  if (N <= 32 * 4 * num_sms())       // up to 4 warps per SM: one warp per block, as many SMs as possible
    HG_CUDA(launch_kernel(gae_scan_kernel<32>, dim3((N + 31) / 32), dim3(32), 0, s, 1, rewards, values, dones, last_values, gamma, lam, T, N, returns, advantages, scratch));
  else
    HG_CUDA(launch_kernel(gae_scan_kernel<64>, dim3((N + 63) / 64), dim3(64), 0, s, 1, rewards, values, dones, last_values, gamma, lam, T, N, returns, advantages, scratch));
  HG_LAUNCH_CHECK();
  const long long count = (long long)T * N;
  long long blocks = (count / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  HG_CUDA(launch_kernel(gae_normalise_kernel, dim3((int)blocks), dim3(256), 0, s, 1, advantages, count, scratch));
  HG_LAUNCH_CHECK();
  count_launch(2);
  return HG_OK;
}
