// The tensor math around the reward / observation kernel of post_physics_step:
//   command_step_kernel  : `_post_physics_step_callback` (humanoid_env.py:1001-1016) with `_resample_commands`
//                          (:1018-1032) and `_push_robots` (:665-681); also the `_resample_commands(env_ids)` of
//                          `reset_idx` (:1127)
//   episode_means_kernel : extras["episode"]["rew_*"] of `reset_idx` (:1141-1143)
// Compiled with -fmad=false: every +,-,* rounds separately in the reference's order, so the resampled commands and
// the pushes are bit-identical to torch-CPU for the same uniform draws (only atan2f differs from libm, <= 2 ulp).
#include "hg_common.cuh"

namespace hg {

constexpr float AUX_TWO_PI_F = 6.2831855f;  // fp32(2*pi)
constexpr float AUX_PI_F = 3.1415927f;

// Philox4x32-10 -> one U[0,1) draw with 24 bits (like torch.rand), keyed by (seed, counter, env, stream word j)
__device__ __forceinline__ float aux_philox_uniform(uint64_t seed, uint64_t counter, uint32_t row, uint32_t j) {
  uint4 ctr = make_uint4(row, 0xC0DE00u + j, (uint32_t)counter, (uint32_t)(counter >> 32));
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return (ctr.x >> 8) * 5.9604644775390625e-8f;
}

// One thread per env; everything an env touches is a few dozen bytes, so the launch is latency-bound by design.
__global__ void command_step_kernel(const __grid_constant__ HgCommandArgs A) {
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.num_envs) return;
  float* cmd = A.commands + (size_t)n * 4;
  float c0 = cmd[0], c1 = cmd[1], c2 = cmd[2], c3 = cmd[3];

  bool resample;
  if (A.mode == HG_CMD_CALLBACK) {
    const long long ep = A.episode_length_buf[n] + A.ep_increment;             // :1005 (after the increment of :780)
    resample = (ep % A.resampling_steps) == 0;
  } else {
    resample = A.reset_buf[n] != 0;                                            // reset_idx(env_ids) :1127
  }
  if (resample) {                                                              // _resample_commands :1024-1032
    float u[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      u[j] = A.u_cmd ? __ldg(A.u_cmd + (size_t)n * 3 + j) : aux_philox_uniform(A.seed, A.counter, (uint32_t)n, j + 8 * A.mode);
    c0 = A.cmd_scale[0] * u[0] + A.cmd_lo[0];                                  // torch_rand_float: (hi - lo) * rand + lo
    c1 = A.cmd_scale[1] * u[1] + A.cmd_lo[1];
    if (A.heading_command) c3 = A.cmd_scale[3] * u[2] + A.cmd_lo[3];
    else c2 = A.cmd_scale[2] * u[2] + A.cmd_lo[2];
    const float keep = sqrtf(c0 * c0 + c1 * c1) > 0.2f ? 1.f : 0.f;            // small commands -> 0 (:1032)
    c0 *= keep;
    c1 *= keep;
  }

  if (A.mode == HG_CMD_CALLBACK) {
    float* root = A.root_states + (size_t)n * 13;
    if (A.heading_command) {                                                   // :1007-1010
      // isaacgym quat_apply(q, b): b + w*t + q_xyz x t, t = 2*(q_xyz x b), with b = forward_vec = (1, 0, 0)
      const float qx = root[3], qy = root[4], qz = root[5], qw = root[6];
      const float bx = 1.f, by = 0.f, bz = 0.f;
      const float tx = (qy * bz - qz * by) * 2.f, ty = (qz * bx - qx * bz) * 2.f, tz = (qx * by - qy * bx) * 2.f;
      const float fx = (bx + qw * tx) + (qy * tz - qz * ty);
      const float fy = (by + qw * ty) + (qz * tx - qx * tz);
      const float heading = atan2f(fy, fx);
      const float a = c3 - heading;
      float m = fmodf(a, AUX_TWO_PI_F);                                        // wrap_to_pi, utils/math.py:46-49
      if (m != 0.f && m < 0.f) m += AUX_TWO_PI_F;                              // torch.remainder (python modulo)
      if (m > AUX_PI_F) m = m - AUX_TWO_PI_F;
      c2 = fminf(fmaxf(0.5f * m, -1.f), 1.f);
    }
    if (A.push) {                                                              // _push_robots :665-681
      float u[5];
#pragma unroll
      for (int j = 0; j < 5; ++j)
        u[j] = A.u_push ? __ldg(A.u_push + (size_t)n * 5 + j) : aux_philox_uniform(A.seed, A.counter, (uint32_t)n, 3 + j);
      if (A.root_vel_prepush) {
#pragma unroll
        for (int j = 0; j < 6; ++j) A.root_vel_prepush[(size_t)n * 6 + j] = root[7 + j];
      }
      const float f0 = A.push_vel_scale * u[0] + A.push_vel_lo, f1 = A.push_vel_scale * u[1] + A.push_vel_lo;
      A.rand_push_force[(size_t)n * 3 + 0] = f0;
      A.rand_push_force[(size_t)n * 3 + 1] = f1;
      root[7] = f0;
      root[8] = f1;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float t = A.push_ang_scale * u[2 + j] + A.push_ang_lo;
        A.rand_push_torque[(size_t)n * 3 + j] = t;
        root[10 + j] = t;
      }
    }
  }
  cmd[0] = c0; cmd[1] = c1; cmd[2] = c2; cmd[3] = c3;
}

// One block per reward term: masked sum over the envs in a fixed order (thread-strided partials in double, then a
// shared-memory tree), so the result is reproducible run to run.
constexpr int MEANS_THREADS = 256;
__global__ void __launch_bounds__(MEANS_THREADS) episode_means_kernel(const float* __restrict__ sums, const uint8_t* __restrict__ reset,
                                                                      int N, float max_len_s, float* __restrict__ means,
                                                                      int32_t* __restrict__ count_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ double s_sum[MEANS_THREADS];
  __shared__ int s_cnt[MEANS_THREADS];
  const int k = blockIdx.x;
  const float* row = sums + (size_t)k * N;
  double acc = 0.0;
  int cnt = 0;
  for (int n = threadIdx.x; n < N; n += MEANS_THREADS) {
    if (reset[n]) {
      acc += (double)row[n];
      ++cnt;
    }
  }
  s_sum[threadIdx.x] = acc;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int off = MEANS_THREADS / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + off];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (s_cnt[0] > 0) means[k] = (float)(s_sum[0] / (double)s_cnt[0]) / max_len_s;     // torch.mean(...) / max_episode_length_s
    if (k == 0 && count_out) count_out[0] = s_cnt[0];
  }
}

}  // namespace hg

extern "C" int hg_command_step(const HgCommandArgs* a, void* stream) {
  using namespace hg;
  HG_CHECK(a != nullptr && a->num_envs > 0 && a->commands != nullptr, "hg_command_step: bad arguments");
  HG_CHECK(a->mode == HG_CMD_CALLBACK || a->mode == HG_CMD_RESET, "hg_command_step: unknown mode %d", a->mode);
  if (a->mode == HG_CMD_CALLBACK) {
    HG_CHECK(a->episode_length_buf != nullptr && a->resampling_steps > 0, "hg_command_step: callback mode needs episode_length_buf and resampling_steps > 0");
    HG_CHECK(a->root_states != nullptr || (!a->heading_command && !a->push), "hg_command_step: heading / push need root_states");
    HG_CHECK(!a->push || (a->rand_push_force && a->rand_push_torque), "hg_command_step: a push needs rand_push_force and rand_push_torque");
  } else {
    HG_CHECK(a->reset_buf != nullptr, "hg_command_step: reset mode needs reset_buf");
  }
  HG_CUDA(launch_kernel(command_step_kernel, dim3((a->num_envs + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 1, *a));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_episode_means(const float* episode_sums, const uint8_t* reset_buf, int num_terms, int num_envs,
                                float max_episode_length_s, float* means_out, int32_t* count_out, void* stream) {
  using namespace hg;
  HG_CHECK(episode_sums && reset_buf && means_out && num_terms > 0 && num_envs > 0 && max_episode_length_s > 0.f,
           "hg_episode_means: bad arguments");
  HG_CUDA(launch_kernel(episode_means_kernel, dim3(num_terms), dim3(MEANS_THREADS), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                        episode_sums, reset_buf, num_envs, max_episode_length_s, means_out, count_out));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}
