// Fused env-step tensor math: one launch replaces the several hundred ATen launches of
// XBotLFreeEnv.post_physics_step (reference humanoid/envs/custom/humanoid_env.py:780-806 and callees).
//
// One block = 16 (or 32) envs.  All per-env inputs and persistent state of the block are first staged into shared memory by
// one table-driven cp.async loop (a single global round trip); the scalar math then runs one THREAD per env and role out of
// shared memory (32x fewer warp instructions than a warp per env), split over "role" warps (observation frames: phase /
// joints and base frame / euler; joint rewards; feet rewards; base-motion rewards) so each env's dependent-latency chain is
// short; the new frames are assembled in shared memory and leave with coalesced stores.  Two history formats:
//   * stacked (generic API, `obs_prev` -> `obs_next`): 16 envs per 512-thread block; warps 4-15 stream the history
//     shift obs_next[n,:] = clip([obs_prev[n,So:], new_frame]) with coalesced 128-byte accesses while warps 0-3 do the
//     math (15 x So floats written, 14 x So read per env; same for the 3-frame critic stack).
//     Algorithmic HBM bytes per env-step (D=12): 4*(47*29 + 73*5) = 6,912 B history + ~0.9 KB state.
//   * history ring (`ring_obs` / `ring_priv`): 16 or 32 envs per 256-thread block (warp 4 takes the
//     base-frame part of the frames, warps 5-7 draw the observation noise while the role warps work); only the NEWEST frame
//     is written (fp32 live + archive, bf16 live + archive), nothing is shifted or re-read; a reset zeroes the env's live
//     frames in place.  Algorithmic bytes per env-step (D=12): 2*4*(47+73) + 2*2*(48+80) = 1,472 B frames + ~0.9 KB state.
//
// Compiled with -fmad=false: every +,-,*,/ is a separately rounded IEEE op in the reference's order, so all
// threshold decisions (contacts, resets, dead-bands, feet clearance) are bit-identical to torch-CPU; only
// sin/cos/exp/atan2/asin and the order of the short DOF sums differ (<= a few ulp).
#include "hg_common.cuh"

namespace hg {

constexpr int FRAME_B_WARP = 4;         // history-ring kernel (8 warps): base-frame part of the observation frames
constexpr int ENV_ROLE_WARPS = 4;       // warps 0-3: per-env math by role; the other warps (stacked format only): history streaming
constexpr int MAX_FRAME = 160;          // >= So, Sp for D <= 32
constexpr float TWO_PI_F = 6.2831855f;  // fp32(2*pi)
constexpr float PI_F = 3.1415927f;

struct Quat { float x, y, z, w; };
struct Vec3 { float x, y, z; };

// isaacgym.torch_utils.quat_rotate_inverse: v*(2w^2-1) - 2w*(q x v) + 2*q*(q.v)
__device__ __forceinline__ Vec3 quat_rotate_inverse(const Quat& q, const Vec3& v) {
  const float s = 2.0f * (q.w * q.w) - 1.0f;
  const Vec3 a = {v.x * s, v.y * s, v.z * s};
  const Vec3 cr = {q.y * v.z - q.z * v.y, q.z * v.x - q.x * v.z, q.x * v.y - q.y * v.x};
  const Vec3 b = {cr.x * q.w * 2.0f, cr.y * q.w * 2.0f, cr.z * q.w * 2.0f};
  const float dot = q.x * v.x + q.y * v.y + q.z * v.z;
  const Vec3 c = {q.x * dot * 2.0f, q.y * dot * 2.0f, q.z * dot * 2.0f};
  return {a.x - b.x + c.x, a.y - b.y + c.y, a.z - b.z + c.z};
}

// The env kernel is instruction-FETCH bound (ncu: 47 % of its stall samples are "no instruction": ~87 KB of straight-line
// code run once per launch by one or two warps per scheduler), so the bulky helpers below exist once (__noinline__) instead
// of once per call site, and the short runtime-bounded loops are not unrolled.
__device__ __noinline__ float wrap_2pi(float a) {   // python-style a % (2*pi), then > pi -> -= 2*pi
  float m = fmodf(a, TWO_PI_F);
  if (m < 0.f) m += TWO_PI_F;
  if (m > PI_F) m -= TWO_PI_F;
  return m;
}

// isaacgym get_euler_xyz + the wrap of get_euler_xyz_tensor (humanoid_env.py:51-56)
__device__ __noinline__ Vec3 euler_xyz_wrapped(const Quat& q) {
  const float sinr = 2.0f * (q.w * q.x + q.y * q.z);
  const float cosr = q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z;
  const float roll = atan2f(sinr, cosr);
  const float sinp = 2.0f * (q.w * q.y - q.z * q.x);
  const float pitch = fabsf(sinp) >= 1.f ? copysignf(1.5707964f, sinp) : asinf(sinp);
  const float siny = 2.0f * (q.w * q.z + q.x * q.y);
  const float cosy = q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z;
  const float yaw = atan2f(siny, cosy);
  return {wrap_2pi(roll), wrap_2pi(pitch), wrap_2pi(yaw)};
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__device__ __noinline__ float sin_2pi(float phase) { return sinf(TWO_PI_F * phase); }
__device__ __forceinline__ float gait_sin(const HgEnvConfig& c, long long ep_len) {
  const float t = (float)ep_len * c.dt;
  const float phase = c.phase_mul_reciprocal ? t * (1.0f / c.cycle_time) : t / c.cycle_time;   // :692
  return sin_2pi(phase);
}
// number of the listed bodies whose contact force exceeds `thr` (check_termination :811-816, _reward_collision :1419-1425)
__device__ __noinline__ int contacts_above(const float* cf, const int* bodies, int n, float thr) {
  int cnt = 0;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const float* f = cf + bodies[i] * 3;
    const float x = f[0], y = f[1], z = f[2];
    cnt += (sqrtf(x * x + y * y + z * z) > thr) ? 1 : 0;
  }
  return cnt;
}
__device__ __forceinline__ float gait_phase(const HgEnvConfig& c, long long ep_len) {
  const float t = (float)ep_len * c.dt;
  return c.phase_mul_reciprocal ? t * (1.0f / c.cycle_time) : t / c.cycle_time;
}

// Philox4x32-10 normal (same generator as the policy head, different stream word)
__device__ __forceinline__ float env_philox_normal(uint64_t seed, uint64_t counter, uint32_t row, uint32_t j) {
  uint4 ctr = make_uint4(row, j, (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0x5EEDu);
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  const float u1 = (ctr.x + 1.0f) * 2.3283064365386963e-10f;
  const float u2 = ctr.y * 2.3283064365386963e-10f;
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// four normals of (seed, counter, row, group): both Box-Muller outputs of both word pairs of one Philox call
__device__ __forceinline__ void env_philox_normal4(uint64_t seed, uint64_t counter, uint32_t row, uint32_t group, float (&out)[4]) {
  uint4 ctr = make_uint4(row, group, (uint32_t)counter, (uint32_t)(counter >> 32) ^ 0x5EEDu);
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll 1
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  const float ra = sqrtf(-2.0f * logf((ctr.x + 1.0f) * 2.3283064365386963e-10f));
  const float rb = sqrtf(-2.0f * logf((ctr.z + 1.0f) * 2.3283064365386963e-10f));
  float sa, ca, sb, cb;
  sincospif(2.0f * (ctr.y * 2.3283064365386963e-10f), &sa, &ca);
  sincospif(2.0f * (ctr.w * 2.3283064365386963e-10f), &sb, &cb);
  out[0] = ra * ca; out[1] = ra * sa; out[2] = rb * cb; out[3] = rb * sb;
}

__device__ __forceinline__ float env_philox_uniform(uint64_t seed, uint64_t counter, uint32_t row) {
  uint4 ctr = make_uint4(row, 0x0DE1A7u, (uint32_t)counter, (uint32_t)(counter >> 32));
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return (ctr.x >> 8) * 5.9604644775390625e-8f;   // [0,1) with 24 bits, like torch.rand
}

// ---- staging: every per-env input / state array of the block's envs is a CONTIGUOUS range of global memory ([E][w]
// floats starting at n0 * w), copied into shared memory with cp.async -- all copies of the block are in flight at once (one
// global round trip for the whole block instead of one per dependent per-thread load); the role warps then read shared
// memory only.  The segment table (source, floats per env, shared-memory offset) is built on the host and handed over as a
// kernel parameter, so the device side is ONE small loop: this kernel runs its code exactly once per launch with one or two
// warps per scheduler, i.e. it pays for every instruction it has to FETCH (the fully inlined version -- one copy loop per
// array -- was 11.2 k instructions = 179 KB of code and 30 % slower than the 5.9 k-instruction kernel it replaced).
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

enum EnvSeg {
  SEG_ROOT = 0, SEG_DOF, SEG_CF, SEG_CMD, SEG_TQ, SEG_ACT, SEG_LA, SEG_LLA, SEG_LDV, SEG_REF, SEG_LRV, SEG_PRE, SEG_PF, SEG_PT,
  SEG_FRIC, SEG_MASS, SEG_AIR, SEG_FH, SEG_LFZ, SEG_EPL, SEG_HIST, SEG_SVAL, SEG_SUMS /* 22 segments: one per reward term */,
  SEG_COUNT = SEG_SUMS + HG_NUM_REWARDS
};
struct EnvStage {
  const float* src[SEG_COUNT];   // nullptr: segment not staged
  int width[SEG_COUNT];          // floats per env
  int dst[SEG_COUNT];            // float offset in shared memory (multiple of 4)
  // produced in shared memory by the kernel itself
  int fobs, fpriv, terms, s2, reset, ep, tout, rg, lc, nsv, ddp, noise, total;
};

// Per-env quantities every role needs (cheap, recomputed by each role's warp instead of being exchanged).
struct EnvBase {
  long long ep;        // episode length after the increment of :780 (REWARD phase) or as stored (OBS-only launch)
  bool reset, time_out;
  float st0, st1;      // stance mask of the pre-reset phase (rewards)
  bool ct0, ct1;       // foot contacts (Fz > 5)
};

// cf: the env's contact forces [NB][3] (shared memory); ep_stored: its episode_length_buf entry
__device__ __forceinline__ EnvBase env_base(const HgEnvConfig& C, const HgEnvBuffers& B, int n, bool doA, const float* cf, long long ep_stored) {
  EnvBase e;
  e.ep = ep_stored + (doA ? 1 : 0);
  if (doA) {                                                                    // check_termination :811-816
    const bool term = contacts_above(cf, C.term_bodies, C.num_term_bodies, 1.0f) > 0;
    e.time_out = (float)e.ep > C.max_episode_length;
    e.reset = term || e.time_out;
  } else {
    e.reset = B.reset_buf[n] != 0;       // decided by an earlier HG_ENV_PHASE_REWARD launch
    e.time_out = B.time_out_buf[n] != 0;
  }
  const float s = gait_sin(C, e.ep);                                            // _get_gait_phase :695-715
  e.st0 = s >= 0.f ? 1.f : 0.f;
  e.st1 = s < 0.f ? 1.f : 0.f;
  if (fabsf(s) < 0.1f) { e.st0 = 1.f; e.st1 = 1.f; }
  e.ct0 = cf[C.feet[0] * 3 + 2] > 5.0f;
  e.ct1 = cf[C.feet[1] * 3 + 2] > 5.0f;
  return e;
}

struct BaseFrame { Vec3 blv, bav, pg, eul; Vec3 lin, ang; float root_z; };

// root: the env's root state [13]; pre: its pre-push root velocities [6] or nullptr
__device__ __forceinline__ BaseFrame base_frame(const float* root, const float* pre, bool need_euler) {
  const Quat q = {root[3], root[4], root[5], root[6]};
  BaseFrame f;
  f.lin = {root[7], root[8], root[9]};
  f.ang = {root[10], root[11], root[12]};
  f.root_z = root[2];
  if (pre) {     // a push this step: the base-frame velocities were taken BEFORE it (:785-786 precede :791)
    f.blv = quat_rotate_inverse(q, Vec3{pre[0], pre[1], pre[2]});
    f.bav = quat_rotate_inverse(q, Vec3{pre[3], pre[4], pre[5]});
  } else {
    f.blv = quat_rotate_inverse(q, f.lin);                                      // :785-787
    f.bav = quat_rotate_inverse(q, f.ang);
  }
  f.pg = quat_rotate_inverse(q, Vec3{0.f, 0.f, -1.f});
  f.eul = need_euler ? euler_xyz_wrapped(q) : Vec3{0.f, 0.f, 0.f};              // :788
  return f;
}

// dst[c] = clip(reset ? 0 : src[c]) for begin <= c < min(count, begin + DEPTH*32), one warp.  Every load of the unit is
// issued before the first store (DEPTH independent 128-byte requests per lane), so a unit costs ONE global round trip.
template <int DEPTH>
__device__ __forceinline__ void shift_load(const float* __restrict__ src, int begin, int count, bool reset, int lane, float (&v)[DEPTH]) {
#pragma unroll
  for (int u = 0; u < DEPTH; ++u) {
    const int c = begin + lane + u * 32;
    v[u] = (!reset && c < count) ? __ldg(src + c) : 0.f;
  }
}
template <int DEPTH>
__device__ __forceinline__ void shift_store(float* __restrict__ dst, __nv_bfloat16* __restrict__ dst16, int begin, int count,
                                            float clip, int lane, const float (&v)[DEPTH]) {
#pragma unroll
  for (int u = 0; u < DEPTH; ++u) {
    const int c = begin + lane + u * 32;
    if (c < count) {
      const float w = clampf(v[u], -clip, clip);
      dst[c] = w;
      if (dst16) dst16[c] = __float2bfloat16_rn(w);
    }
  }
}
constexpr int SHIFT_UNIT = 672;   // floats per streaming unit (21 loads per lane): the 658-float actor history of D = 12 is one unit

// One block = E envs.  (1) all threads stage the envs' inputs and persistent state into shared memory (cp.async, one round
// trip); (2) one THREAD per env and role does the scalar math out of shared memory -- the block's first four warps split the
// work by role so the dependent chain per env is ~4x shorter:
//   warp 0: observation frames (phase, euler, base-frame velocities, joint terms)      compute_observations :818-887
//   warp 1: joint-space rewards     warp 2: feet / contact rewards (+ their state)     warp 3: base-motion rewards
// (3) all threads: episode sums, hand-over, noise + clip of the newest frame (flat over [env, column]) and its stores.
template <int E, int ENV_THREADS, bool RING>
__global__ void __launch_bounds__(ENV_THREADS, RING ? 4 : 2) env_step_kernel(const __grid_constant__ HgEnvConfig C,
                                                                            const __grid_constant__ HgEnvBuffers B,
                                                                            const __grid_constant__ EnvStage L) {
  extern __shared__ __align__(16) float s_env[];
  if (threadIdx.x == 0) trace(0x100);
  const int D = C.D, NB = C.num_bodies, off = C.leg_offset;
  const int So = 11 + 3 * D, Sp = 25 + 4 * D;
  float* s_fobs = s_env + L.fobs;                     // [E][So]
  float* s_fpriv = s_env + L.fpriv;                   // [E][Sp]
  float* s_terms = s_env + L.terms;                   // [HG_NUM_REWARDS][E]
  float* s_s2 = s_env + L.s2;                         // [E] sin of the post-reset phase
  int* s_reset = reinterpret_cast<int*>(s_env + L.reset);
  int* s_ep = reinterpret_cast<int*>(s_env + L.ep);   // episode length to store
  int* s_tout = reinterpret_cast<int*>(s_env + L.tout);
  const float* s_nsv = s_env + L.nsv;
  const float* s_ddp = s_env + L.ddp;
  const float* s_sums = s_env + L.dst[SEG_SUMS];      // [HG_NUM_REWARDS][E] episode sums before this step

  const int tid = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * E;
  const int envs_here = min(E, C.num_envs - n0);
  const int n = n0 + lane;
  const bool ok = lane < envs_here;
  const bool doA = (C.phases & HG_ENV_PHASE_REWARD) != 0, doB = (C.phases & HG_ENV_PHASE_OBS) != 0;

  pdl_wait();
  if (threadIdx.x == 0) trace(0x101);
  // ---------------------------------------------------------------- (1) stage: one warp per segment
#pragma unroll 1
  for (int sgi = warp; sgi < SEG_COUNT; sgi += ENV_THREADS / 32) {
    const float* src = L.src[sgi];
    if (src == nullptr) continue;
    const int w = L.width[sgi], count = envs_here * w;
    src += (size_t)n0 * w;
    const uint32_t d = smem_u32(s_env + L.dst[sgi]);
    if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (count & 3) == 0) {
#pragma unroll 1
      for (int i = lane; i < (count >> 2); i += 32) cp_async16(d + 16 * i, src + 4 * i);
    } else {
#pragma unroll 1
      for (int i = lane; i < count; i += 32) cp_async4(d + 4 * i, src + i);
    }
  }
#pragma unroll 1
  for (int i = tid; i < envs_here * 52; i += ENV_THREADS) {            // rigid-body rows of the two feet and the two knees only
    const int e = i / 52, r = (i - e * 52) / 13, c = i - e * 52 - r * 13;
    const bool cr = C.num_rigid_bodies > 0;          // compact rigid-body table with its own indices
    const int body = r == 0 ? (cr ? C.rigid_feet[0] : C.feet[0]) : r == 1 ? (cr ? C.rigid_feet[1] : C.feet[1])
                            : r == 2 ? (cr ? C.rigid_knees[0] : C.knees[0]) : (cr ? C.rigid_knees[1] : C.knees[1]);
    cp_async4(smem_u32(s_env + L.rg + i), B.rigid_state + ((size_t)(n0 + e) * (cr ? C.num_rigid_bodies : NB) + body) * 13 + c);
  }
  if (tid < So) cp_async4(smem_u32(s_env + L.nsv + tid), B.noise_scale_vec + tid);
  if (tid < D) cp_async4(smem_u32(s_env + L.ddp + tid), B.default_dof_pos + tid);
  if (tid < envs_here) {                                // last_contacts: two bytes per env
    const uint8_t* lcp = B.last_contacts + (size_t)(n0 + tid) * 2;
    reinterpret_cast<int*>(s_env + L.lc)[tid] = (int)lcp[0] | ((int)lcp[1] << 8);
  }
  cp_async_wait_all();
  __syncthreads();
  if (threadIdx.x == 0) trace(0x105);

  const float* cf = s_env + L.dst[SEG_CF] + lane * NB * 3;
  const float* rg = s_env + L.rg + lane * 52;         // rows: foot 0, foot 1, knee 0, knee 1
  const float* dof = s_env + L.dst[SEG_DOF] + lane * 2 * D;
  const float* root = s_env + L.dst[SEG_ROOT] + lane * 13;
  const float* pre = B.root_vel_prepush ? s_env + L.dst[SEG_PRE] + lane * 6 : nullptr;
  const float* cmd = s_env + L.dst[SEG_CMD] + lane * 4;
  const float* a_in = s_env + L.dst[SEG_ACT] + lane * D;
  const float* la_in = s_env + L.dst[SEG_LA] + lane * D;
  const float* lla_in = s_env + L.dst[SEG_LLA] + lane * D;
  const bool clip_a = B.actions_in != nullptr;
  auto action_of = [&](const float* row, int i) { return clip_a ? clampf(row[i], -B.clip_actions, B.clip_actions) : row[i]; };
  auto ep_stored = [&](int el) { return reinterpret_cast<const long long*>(s_env + L.dst[SEG_EPL])[el]; };
  float* T = s_terms + lane;                          // T[k * E]
#define HG_TERM(k, val) T[(k) * E] = (val) * C.reward_scale_dt[k]

  // N(0,1) draws of the newest actor frame, [E][So rounded up to 4]: they depend on nothing but (seed, step, env, column
  // group), so in the ring kernel the three warps without a role produce them WHILE the role warps do the per-env math
  // (the stacked kernel's spare warps stream the history instead: there all threads draw them after the barrier)
  float* s_noise = s_env + L.noise;
  const int ngroups = (So + 3) >> 2, So4 = ngroups * 4;
  auto draw_noise = [&](int first, int stride) {
    if (C.noise_level == 0.f) return;
#pragma unroll 1
    for (int i = first; i < envs_here * ngroups; i += stride) {
      const int el = i / ngroups, g = i - el * ngroups;
      const int ne = n0 + el, j0 = 4 * g, cols = min(4, So - j0);
      float e4[4] = {0.f, 0.f, 0.f, 0.f};
      if (B.noise) {
        for (int k = 0; k < cols; ++k) e4[k] = __ldg(B.noise + (size_t)ne * So + j0 + k);
      } else {
        bool any = false;
#pragma unroll 1
        for (int k = 0; k < cols; ++k) any = any || s_nsv[j0 + k] != 0.f;
        if (any) env_philox_normal4(B.seed, B.counter, ne, g, e4);     // one Philox4x32-10 call: four normals
      }
      *reinterpret_cast<float4*>(s_noise + el * So4 + j0) = make_float4(e4[0], e4[1], e4[2], e4[3]);
    }
  };
  if (RING && warp > FRAME_B_WARP) draw_noise(tid - 32 * (FRAME_B_WARP + 1), ENV_THREADS - 32 * (FRAME_B_WARP + 1));

  if (!RING && warp >= ENV_ROLE_WARPS) {
    // ---------------------------------------------------------------- history shift (:883-887): needs only the reset flag,
    // so these 12 warps stream the 14 old actor frames / 2 old critic frames while warps 0-3 do the per-env math.
    if (doB) {
      // Work units: the actor history of an env in slabs of SHIFT_UNIT floats, then its critic history in slabs of
      // SHIFT_UNIT; units are dealt round-robin to the 12 streaming warps (16 envs x 2 units at D = 12: <= 3 round trips
      // each, at most 2 of them long).
      const int cnt_o = (C.frame_stack - 1) * So, cnt_p = (C.c_frame_stack - 1) * Sp;
      const int upo = (cnt_o + SHIFT_UNIT - 1) / SHIFT_UNIT, upe = upo + (cnt_p + SHIFT_UNIT - 1) / SHIFT_UNIT;
      for (int u = warp - ENV_ROLE_WARPS; u < envs_here * upe; u += ENV_THREADS / 32 - ENV_ROLE_WARPS) {
        const int part = u / envs_here, el = u - part * envs_here;   // part-major: every warp gets a mix of long and short units
        const int ne = n0 + el;
        const bool reset = env_base(C, B, ne, doA, s_env + L.dst[SEG_CF] + el * NB * 3, ep_stored(el)).reset;       // warp-uniform
        const bool actor = part < upo;
        const float* src = actor ? B.obs_prev + (size_t)ne * B.obs_prev_ld + So : B.priv_prev + (size_t)ne * B.priv_prev_ld + Sp;
        float* dst = actor ? B.obs_next + (size_t)ne * B.obs_next_ld : B.priv_next + (size_t)ne * B.priv_next_ld;
        __nv_bfloat16* d16 = nullptr;
        if (actor && B.obs16_next) d16 = reinterpret_cast<__nv_bfloat16*>(B.obs16_next) + (size_t)ne * B.obs16_next_ld;
        if (!actor && B.priv16_next) d16 = reinterpret_cast<__nv_bfloat16*>(B.priv16_next) + (size_t)ne * B.priv16_next_ld;
        const int begin = (actor ? part : part - upo) * SHIFT_UNIT, count = actor ? cnt_o : cnt_p;
        float v[SHIFT_UNIT / 32];
        shift_load(src, begin, count, reset, lane, v);
        shift_store(dst, d16, begin, count, C.clip_obs, lane, v);
      }
    }
  } else if (ok && (warp < ENV_ROLE_WARPS || warp == FRAME_B_WARP)) {
    // (the stacked format has no spare warp: its warp 0 also does the base-frame part)
    const bool frame_a = warp == 0, frame_b = RING ? warp == FRAME_B_WARP : warp == 0;
    const EnvBase e = env_base(C, B, n, doA, cf, ep_stored(lane));
    if (frame_b && doB) {
      // ------------------------------------------------------------ observation frames, base-frame part (its own warp: the
      // quaternion rotations + euler angles are the longest dependent chain of the frame)
      const BaseFrame f = base_frame(root, pre, true);
      float* fo = s_fobs + lane * So;
      float* t = s_fpriv + lane * Sp + 5 + 4 * D;
      fo[5 + 3 * D + 0] = f.bav.x * C.ang_vel_scale; fo[5 + 3 * D + 1] = f.bav.y * C.ang_vel_scale; fo[5 + 3 * D + 2] = f.bav.z * C.ang_vel_scale;
      fo[8 + 3 * D + 0] = f.eul.x * C.quat_scale; fo[8 + 3 * D + 1] = f.eul.y * C.quat_scale; fo[8 + 3 * D + 2] = f.eul.z * C.quat_scale;
      t[0] = f.blv.x * C.lin_vel_scale; t[1] = f.blv.y * C.lin_vel_scale; t[2] = f.blv.z * C.lin_vel_scale;
      t[3] = f.bav.x * C.ang_vel_scale; t[4] = f.bav.y * C.ang_vel_scale; t[5] = f.bav.z * C.ang_vel_scale;
      t[6] = f.eul.x * C.quat_scale; t[7] = f.eul.y * C.quat_scale; t[8] = f.eul.z * C.quat_scale;
    }
    if (frame_a) {
      // ------------------------------------------------------------ observation frames (post-reset state): phase, commands, joints
      if (doB) {
        const long long ep = e.reset ? 0 : e.ep;                                // reset_idx :1138
        const float ph = gait_phase(C, ep);
        const float s2 = sin_2pi(ph), c2 = cosf(TWO_PI_F * ph);
        float so0 = s2 >= 0.f ? 1.f : 0.f, so1 = s2 < 0.f ? 1.f : 0.f;
        const bool dsup = fabsf(s2) < 0.1f;
        if (dsup) { so0 = 1.f; so1 = 1.f; }
        const float sl = s2 > 0.f ? 0.f : s2, sr = s2 < 0.f ? 0.f : s2;         // compute_ref_state :717-744
        const float sc1 = C.target_joint_pos_scale, sc2 = 2.0f * C.target_joint_pos_scale;
        float* fo = s_fobs + lane * So;
        float* fp = s_fpriv + lane * Sp;
        const float cmd0 = cmd[0], cmd1 = cmd[1], cmd2 = cmd[2];
        fo[0] = s2; fo[1] = c2; fo[2] = cmd0 * C.lin_vel_scale; fo[3] = cmd1 * C.lin_vel_scale; fo[4] = cmd2 * C.ang_vel_scale;
        fp[0] = s2; fp[1] = c2; fp[2] = fo[2]; fp[3] = fo[3]; fp[4] = fo[4];
#pragma unroll 1
        for (int i = 0; i < D; ++i) {
          const float px = dof[2 * i], py = dof[2 * i + 1];
          const float dflt = s_ddp[i];
          const float a = e.reset ? 0.f : action_of(a_in, i);
          const int j = i - off;
          float ref = 0.f;
          if (j == 2 || j == 4) ref = sl * sc1;
          if (j == 3) ref = sl * sc2;
          if (j == 8 || j == 10) ref = sr * sc1;
          if (j == 9) ref = sr * sc2;
          if (dsup) ref = 0.f;
          const float qv = (px - dflt) * C.dof_pos_scale, dqv = py * C.dof_vel_scale;
          fo[5 + i] = qv; fo[5 + D + i] = dqv; fo[5 + 2 * D + i] = a;
          fp[5 + i] = qv; fp[5 + D + i] = dqv; fp[5 + 2 * D + i] = a; fp[5 + 3 * D + i] = px - ref;
        }
        float* t = fp + 5 + 4 * D;
        const float* pf = s_env + L.dst[SEG_PF] + lane * 3;
        const float* pt = s_env + L.dst[SEG_PT] + lane * 3;
        t[9] = pf[0]; t[10] = pf[1];
        t[11] = pt[0]; t[12] = pt[1]; t[13] = pt[2];
        t[14] = s_env[L.dst[SEG_FRIC] + lane];
        t[15] = s_env[L.dst[SEG_MASS] + lane] / 30.0f;
        t[16] = so0; t[17] = so1;
        t[18] = e.ct0 ? 1.f : 0.f; t[19] = e.ct1 ? 1.f : 0.f;
        s_s2[lane] = s2;
        s_ep[lane] = (int)ep;
      } else {
        s_ep[lane] = (int)e.ep;
      }
      s_reset[lane] = e.reset ? 1 : 0;
      s_tout[lane] = e.time_out ? 1 : 0;
    } else if (doA && warp == 1) {
      // ------------------------------------------------------------ joint-space rewards
      const float* tq_in = s_env + L.dst[SEG_TQ] + lane * D;
      const float* ldv_in = s_env + L.dst[SEG_LDV] + lane * D;
      const float* ref_in = s_env + L.dst[SEG_REF] + lane * D;
      float t1 = 0.f, t2 = 0.f, t3 = 0.f, jd_all = 0.f, l01 = 0.f, r01 = 0.f, acc2 = 0.f, vel2 = 0.f, tq2 = 0.f, jp2 = 0.f;
#pragma unroll 1
      for (int i = 0; i < D; ++i) {
        const float px = dof[2 * i], py = dof[2 * i + 1];
        const float a = action_of(a_in, i), la = la_in[i], lla = lla_in[i];
        const float d1 = la - a, d2 = a + lla - 2.0f * la;
        t1 += d1 * d1; t2 += d2 * d2; t3 += fabsf(a);                           // action_smoothness :1427-1437
        const float jd = px - s_ddp[i];                                          // default_joint_pos :1259-1269
        jd_all += jd * jd;
        if (i == off || i == off + 1) l01 += jd * jd;
        if (i == off + 6 || i == off + 7) r01 += jd * jd;
        const float ac = (ldv_in[i] - py) / C.dt;                                // dof_acc :1412-1417
        acc2 += ac * ac;
        vel2 += py * py;                                                         // dof_vel :1405-1410
        const float tq = tq_in[i];                                               // torques :1398-1403
        tq2 += tq * tq;
        const float dj = px - ref_in[i];                                         // joint_pos :1170-1178 (previous step's ref)
        jp2 += dj * dj;
      }
      HG_TERM(HG_R_ACTION_SMOOTHNESS, t1 + t2 + 0.05f * t3);
      float yr = sqrtf(l01) + sqrtf(r01);
      yr = clampf(yr - 0.1f, 0.f, 50.f);
      HG_TERM(HG_R_DEFAULT_JOINT_POS, expf(-yr * 100.0f) - 0.01f * sqrtf(jd_all));
      HG_TERM(HG_R_DOF_ACC, acc2);
      HG_TERM(HG_R_DOF_VEL, vel2);
      HG_TERM(HG_R_TORQUES, tq2);
      const float nrm = sqrtf(jp2);
      HG_TERM(HG_R_JOINT_POS, expf(-2.0f * nrm) - 0.2f * clampf(nrm, 0.f, 0.5f));
    } else if (doA && warp == 2) {
      // ------------------------------------------------------------ feet / contact rewards (+ their persistent state)
      const float foot_z0 = rg[0 * 13 + 2], foot_z1 = rg[1 * 13 + 2];
      {  // base_height :1271-1281
        const float meas = (foot_z0 * e.st0 + foot_z1 * e.st1) / (e.st0 + e.st1);
        const float bh = root[2] - (meas - 0.05f);
        HG_TERM(HG_R_BASE_HEIGHT, expf(-fabsf(bh - C.base_height_target) * 100.0f));
      }
      {  // collision :1419-1425
        HG_TERM(HG_R_COLLISION, (float)contacts_above(cf, C.pen_bodies, C.num_pen_bodies, 0.1f));
      }
      {  // feet_air_time :1216-1230 (stateful)
        float air0 = s_env[L.dst[SEG_AIR] + lane * 2], air1 = s_env[L.dst[SEG_AIR] + lane * 2 + 1];
        const int lcw = reinterpret_cast<const int*>(s_env + L.lc)[lane];
        const bool lc0 = (lcw & 0xFF) != 0, lc1 = (lcw >> 8) != 0;
        const bool cfl0 = e.ct0 || (e.st0 != 0.f) || lc0, cfl1 = e.ct1 || (e.st1 != 0.f) || lc1;
        const float first0 = (air0 > 0.f && cfl0) ? 1.f : 0.f, first1 = (air1 > 0.f && cfl1) ? 1.f : 0.f;
        air0 += C.dt;
        air1 += C.dt;
        HG_TERM(HG_R_FEET_AIR_TIME, clampf(air0, 0.f, 0.5f) * first0 + clampf(air1, 0.f, 0.5f) * first1);
        air0 *= cfl0 ? 0.f : 1.f;
        air1 *= cfl1 ? 0.f : 1.f;
        if (e.reset && doB) { air0 = 0.f; air1 = 0.f; }                          // reset_idx :1137
        *reinterpret_cast<float2*>(B.feet_air_time + (size_t)n * 2) = make_float2(air0, air1);
        *reinterpret_cast<uchar2*>(B.last_contacts + (size_t)n * 2) = make_uchar2(e.ct0 ? 1 : 0, e.ct1 ? 1 : 0);
      }
      {  // feet_clearance :1342-1363 (stateful)
        float fh0 = s_env[L.dst[SEG_FH] + lane * 2], fh1 = s_env[L.dst[SEG_FH] + lane * 2 + 1];
        const float fzz0 = foot_z0 - 0.05f, fzz1 = foot_z1 - 0.05f;
        fh0 += fzz0 - s_env[L.dst[SEG_LFZ] + lane * 2];
        fh1 += fzz1 - s_env[L.dst[SEG_LFZ] + lane * 2 + 1];
        const float p0 = fabsf(fh0 - C.target_feet_height) < 0.01f ? 1.f : 0.f;
        const float p1 = fabsf(fh1 - C.target_feet_height) < 0.01f ? 1.f : 0.f;
        HG_TERM(HG_R_FEET_CLEARANCE, p0 * (1.f - e.st0) + p1 * (1.f - e.st1));
        fh0 *= e.ct0 ? 0.f : 1.f;
        fh1 *= e.ct1 ? 0.f : 1.f;
        *reinterpret_cast<float2*>(B.feet_height + (size_t)n * 2) = make_float2(fh0, fh1);
        *reinterpret_cast<float2*>(B.last_feet_z + (size_t)n * 2) = make_float2(fzz0, fzz1);
      }
      {  // feet_contact_forces :1251-1257, feet_contact_number :1232-1240
        const float* f0 = cf + C.feet[0] * 3;
        const float* f1 = cf + C.feet[1] * 3;
        const float x0 = f0[0], y0 = f0[1], z0 = f0[2], x1 = f1[0], y1 = f1[1], z1 = f1[2];
        const float nf0 = sqrtf(x0 * x0 + y0 * y0 + z0 * z0), nf1 = sqrtf(x1 * x1 + y1 * y1 + z1 * z1);
        HG_TERM(HG_R_FEET_CONTACT_FORCES, clampf(nf0 - C.max_contact_force, 0.f, 400.f) + clampf(nf1 - C.max_contact_force, 0.f, 400.f));
        const float e0 = ((e.ct0 ? 1.f : 0.f) == e.st0) ? 1.f : -0.3f, e1 = ((e.ct1 ? 1.f : 0.f) == e.st1) ? 1.f : -0.3f;
        HG_TERM(HG_R_FEET_CONTACT_NUMBER, (e0 + e1) / 2.0f);
      }
      {  // feet_distance :1180-1190, knee_distance :1192-1202   (rows of rg: foot 0, foot 1, knee 0, knee 1)
        auto dist_reward = [&](int r0, int r1, float max_d) {
          const float dx = rg[r0 * 13] - rg[r1 * 13], dy = rg[r0 * 13 + 1] - rg[r1 * 13 + 1];
          const float d = sqrtf(dx * dx + dy * dy);
          const float d_min = clampf(d - C.min_dist, -0.5f, 0.f), d_max = clampf(d - max_d, 0.f, 0.5f);
          return (expf(-fabsf(d_min) * 100.0f) + expf(-fabsf(d_max) * 100.0f)) / 2.0f;
        };
        HG_TERM(HG_R_FEET_DISTANCE, dist_reward(0, 1, C.max_dist));
        HG_TERM(HG_R_KNEE_DISTANCE, dist_reward(2, 3, C.max_dist / 2.0f));
      }
      {  // foot_slip :1204-1214 (columns 10:12 as written in the reference)
        const float ax = rg[0 * 13 + 10], ay = rg[0 * 13 + 11];
        const float bx = rg[1 * 13 + 10], by = rg[1 * 13 + 11];
        HG_TERM(HG_R_FOOT_SLIP, sqrtf(sqrtf(ax * ax + ay * ay)) * (e.ct0 ? 1.f : 0.f) + sqrtf(sqrtf(bx * bx + by * by)) * (e.ct1 ? 1.f : 0.f));
      }
    } else if (doA && warp == 3) {
      // ------------------------------------------------------------ base-motion rewards
      const BaseFrame f = base_frame(root, pre, true);
      const float cmd0 = cmd[0], cmd1 = cmd[1], cmd2 = cmd[2];
      {  // base_acc :1283-1290
        const float* lrv = s_env + L.dst[SEG_LRV] + lane * 6;
        const float a0 = lrv[0] - f.lin.x, a1 = lrv[1] - f.lin.y, a2 = lrv[2] - f.lin.z;
        const float a3 = lrv[3] - f.ang.x, a4 = lrv[4] - f.ang.y, a5 = lrv[5] - f.ang.z;
        HG_TERM(HG_R_BASE_ACC, expf(-sqrtf(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5) * 3.0f));
      }
      {  // low_speed :1365-1396
        const float av = fabsf(f.blv.x), ac = fabsf(cmd0);
        const bool low = av < 0.5f * ac, high = av > 1.2f * ac;
        float v = 0.f;
        if (low) v = -1.0f;
        if (high) v = 0.f;
        if (!(low || high)) v = 1.2f;
        if (sgnf(f.blv.x) != sgnf(cmd0)) v = -2.0f;
        HG_TERM(HG_R_LOW_SPEED, v * (ac > 0.1f ? 1.f : 0.f));
      }
      HG_TERM(HG_R_ORIENTATION, (expf(-(fabsf(f.eul.x) + fabsf(f.eul.y)) * 10.0f) + expf(-sqrtf(f.pg.x * f.pg.x + f.pg.y * f.pg.y) * 20.0f)) / 2.0f);
      const float ex = cmd0 - f.blv.x, ey = cmd1 - f.blv.y;
      const float e_l = sqrtf(ex * ex + ey * ey), e_a = fabsf(cmd2 - f.bav.z);
      HG_TERM(HG_R_TRACK_VEL_HARD, (expf(-e_l * 10.0f) + expf(-e_a * 10.0f)) / 2.0f - 0.2f * (e_l + e_a));
      const float ea2 = (cmd2 - f.bav.z) * (cmd2 - f.bav.z);
      HG_TERM(HG_R_TRACKING_ANG_VEL, expf(-ea2 * C.tracking_sigma));
      HG_TERM(HG_R_TRACKING_LIN_VEL, expf(-(ex * ex + ey * ey) * C.tracking_sigma));
      HG_TERM(HG_R_VEL_MISMATCH_EXP, (expf(-(f.blv.z * f.blv.z) * 10.0f) + expf(-sqrtf(f.bav.x * f.bav.x + f.bav.y * f.bav.y) * 5.0f)) / 2.0f);
    }
  }
#undef HG_TERM
  if (lane == 0) trace(0x110 + warp);
  __syncthreads();
  // Every block is resident by now (the whole grid triggers only once its last block got here): the dependent launch
  // becomes resident behind this one's stores instead of after the grid has drained.
  pdl_launch_dependents();
  if (threadIdx.x == 0) trace(0x102);

  if (ok && warp == 0) {
    B.episode_length_buf[n] = s_ep[lane];
    if (doA) {
      B.reset_buf[n] = s_reset[lane] ? 1 : 0;
      B.time_out_buf[n] = s_tout[lane] ? 1 : 0;
    }
  }
  // ---------------------------------------------------------------- episode sums (all threads, coalesced over [term, env])
  if (doA) {
#pragma unroll 1
    for (int i = tid; i < HG_NUM_REWARDS * E; i += ENV_THREADS) {
      const int k = i / E, el = i % E;                    // (E is a compile-time power of two: no division)
      if (el < envs_here && C.reward_scale_dt[k] != 0.f) {
        const float term_k = s_terms[k * E + el];
        const size_t idx = (size_t)k * C.num_envs + n0 + el;
        B.episode_sums[idx] = (s_reset[el] && doB) ? 0.f : s_sums[k * E + el] + term_k;   // zeroed by reset_idx (:1144)
        if (B.reward_terms) B.reward_terms[idx] = term_k;
      }
    }
  }
  // ---------------------------------------------------------------- total reward (alphabetical order :894-898), clip
  if (doA && ok && warp == 1) {
    float rew = 0.f;
#pragma unroll 1
    for (int k = 0; k < HG_NUM_REWARDS; ++k)
      if (C.reward_scale_dt[k] != 0.f) rew += s_terms[k * E + lane];
    if (C.only_positive_rewards) rew = fmaxf(rew, 0.f);                              // :899-900
    B.rew_buf[n] = rew;
    if (B.store_rewards) {       // the learner's transition record (same rounding as hg_store_rewards_dones)
      B.store_rewards[n] = __fadd_rn(rew, __fmul_rn(B.store_gamma, __fmul_rn(s_env[L.dst[SEG_SVAL] + lane], s_tout[lane] ? 1.f : 0.f)));
      B.store_dones[n] = s_reset[lane] ? 1 : 0;
    }
  }
  if (threadIdx.x == 0) trace(0x103);
  if (!doB) return;
  if (!doA && ok && warp == 1 && s_reset[lane]) {
#pragma unroll 1
    for (int k = 0; k < HG_NUM_REWARDS; ++k) B.episode_sums[(size_t)k * C.num_envs + n] = 0.f;
    B.feet_air_time[n * 2] = 0.f; B.feet_air_time[n * 2 + 1] = 0.f;
  }

  // ---------------------------------------------------------------- hand-over :802-806 + reset zeroing :1132-1139, coalesced over [env, dof]
  // (the flat phases below walk [env = warp, warp + 8, ..][column = lane, lane + 32, ..]: no runtime integer division, which
  // costs ~100 dependent cycles apiece and used to head every iteration of these short loops)
#pragma unroll 1
  for (int el = warp; el < envs_here; el += ENV_THREADS / 32)
  for (int d = lane; d < D; d += 32) {
    const int i = el * D + d;
    const size_t k = (size_t)(n0 + el) * D + d;
    const bool reset = s_reset[el] != 0;
    const float araw = s_env[L.dst[SEG_ACT] + i];
    const float a = reset ? 0.f : (clip_a ? clampf(araw, -B.clip_actions, B.clip_actions) : araw);
    const float la = reset ? 0.f : s_env[L.dst[SEG_LA] + i];
    B.last_last_actions[k] = la;
    B.last_actions[k] = a;
    if (reset || clip_a) B.actions[k] = a;
    B.last_dof_vel[k] = s_env[L.dst[SEG_DOF] + 2 * i + 1];
    const float s2 = s_s2[el];
    const float sl = s2 > 0.f ? 0.f : s2, sr = s2 < 0.f ? 0.f : s2;
    const float sc1 = C.target_joint_pos_scale, sc2 = 2.0f * C.target_joint_pos_scale;
    const int j = d - off;
    float ref = 0.f;
    if (j == 2 || j == 4) ref = sl * sc1;
    if (j == 3) ref = sl * sc2;
    if (j == 8 || j == 10) ref = sr * sc1;
    if (j == 9) ref = sr * sc2;
    if (fabsf(s2) < 0.1f) ref = 0.f;
    B.ref_dof_pos[k] = ref;
  }
  for (int i = tid; i < envs_here * 6; i += ENV_THREADS) {
    const int el = i / 6, d = i - el * 6;                   // (constant divisor)
    B.last_root_vel[(size_t)n0 * 6 + i] = s_env[L.dst[SEG_ROOT] + el * 13 + 7 + d];
  }
  if (threadIdx.x == 0) trace(0x104);

  // ---------------------------------------------------------------- newest frame: + noise (:877), clip (:654-657), flat over [env, column]
  if (!RING) {
    draw_noise(tid, ENV_THREADS);
    __syncthreads();
  }
#pragma unroll 1
  for (int el = warp; el < envs_here; el += ENV_THREADS / 32) {
#pragma unroll 1
    for (int j = lane; j < So; j += 32) {
      float v = s_fobs[el * So + j];
      if (C.noise_level != 0.f) v = v + s_noise[el * So4 + j] * s_nsv[j] * C.noise_level;
      s_fobs[el * So + j] = clampf(v, -C.clip_obs, C.clip_obs);
    }
#pragma unroll 1
    for (int j = lane; j < Sp; j += 32) s_fpriv[el * Sp + j] = clampf(s_fpriv[el * Sp + j], -C.clip_obs, C.clip_obs);
  }
  __syncwarp();          // the stores below walk the same [env = warp][column = lane] map: a warp reads only what it wrote
  if constexpr (!RING) {
    for (int el = warp; el < envs_here; el += ENV_THREADS / 32)
    for (int j = lane; j < So; j += 32) {
      const int ne = n0 + el;
      const float v = s_fobs[el * So + j];
      B.obs_next[(size_t)ne * B.obs_next_ld + (C.frame_stack - 1) * So + j] = v;
      if (B.obs16_next)
        reinterpret_cast<__nv_bfloat16*>(B.obs16_next)[(size_t)ne * B.obs16_next_ld + (C.frame_stack - 1) * So + j] = __float2bfloat16_rn(v);
    }
    for (int el = warp; el < envs_here; el += ENV_THREADS / 32)
    for (int j = lane; j < Sp; j += 32) {
      const int ne = n0 + el;
      const float v = s_fpriv[el * Sp + j];
      B.priv_next[(size_t)ne * B.priv_next_ld + (C.c_frame_stack - 1) * Sp + j] = v;
      if (B.priv16_next)
        reinterpret_cast<__nv_bfloat16*>(B.priv16_next)[(size_t)ne * B.priv16_next_ld + (C.c_frame_stack - 1) * Sp + j] = __float2bfloat16_rn(v);
    }
  } else {
    // history ring: the newest frame into slot ring_row + frames - 1 of live32 / arch32 (fp32) and live16 / arch16 (bf16 pairs,
    // pad columns zero); a reset env first gets its previous frames - 1 live slots zeroed (reset_idx :1159-1163) -- the
    // archives keep them for the rows of the iteration that still need them (first_valid marks where a row's history starts)
    auto put = [&](const HgHistoryRing& R, const float* frames_s) {
      const int slot = B.ring_row + R.frames - 1;
      const int pairs = R.pitch16 >> 1;
#pragma unroll 1
      for (int el = warp; el < envs_here; el += ENV_THREADS / 32) {      // one warp per env, lanes over the frame's columns
        const size_t env_base = (size_t)(n0 + el) * R.slots;
        const float* fr = frames_s + el * R.width;
        float* l32 = R.live32 + (env_base + slot) * R.width;
        float* a32 = R.arch32 + (env_base + slot) * R.width;
#pragma unroll 1
        for (int j = lane; j < R.width; j += 32) {
          const float v = fr[j];
          l32[j] = v;
          a32[j] = v;
        }
        uint32_t* l16 = reinterpret_cast<uint32_t*>(R.live16) + (env_base + slot) * pairs;
        uint32_t* a16 = reinterpret_cast<uint32_t*>(R.arch16) + (env_base + slot) * pairs;
#pragma unroll 1
        for (int j = lane; j < pairs; j += 32) {
          const float lo = 2 * j < R.width ? fr[2 * j] : 0.f, hi = 2 * j + 1 < R.width ? fr[2 * j + 1] : 0.f;
          const uint32_t w = pack_bf16x2(lo, hi);
          l16[j] = w;
          a16[j] = w;
        }
        if (s_reset[el]) {       // (rare) a reset env: zero its live history
          float* z32 = R.live32 + (env_base + B.ring_row) * R.width;
#pragma unroll 1
          for (int j = lane; j < (R.frames - 1) * R.width; j += 32) z32[j] = 0.f;
          uint32_t* z16 = reinterpret_cast<uint32_t*>(R.live16) + (env_base + B.ring_row) * pairs;
#pragma unroll 1
          for (int j = lane; j < (R.frames - 1) * pairs; j += 32) z16[j] = 0u;
        }
      }
    };
    put(B.ring_obs, s_fobs);
    put(B.ring_priv, s_fpriv);
    if (tid < envs_here) {
      const int ne = n0 + tid;
      const int hs = s_reset[tid] ? B.ring_row : reinterpret_cast<const int*>(s_env + L.dst[SEG_HIST])[tid];
      B.ring_hist_start[ne] = hs;
      B.ring_first_valid[(size_t)B.ring_row * C.num_envs + ne] = hs;
    }
  }
  if (lane == 0) trace(0x120 + warp);
}

HG_TRACE_SETTER(trace_set_env)

__global__ void action_torque_kernel(const __grid_constant__ HgActionArgs P) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)P.num_envs * P.D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / P.D), d = (int)(i - (long long)n * P.D);
    float a = P.actions[i];
    if (P.raw_actions) {
      const float delay = P.delay ? __ldg(P.delay + n) : env_philox_uniform(P.seed, P.counter, (uint32_t)n);
      a = (1.0f - delay) * __ldg(P.raw_actions + i) + delay * a;                          // :629
      const float e = P.noise ? __ldg(P.noise + i) : env_philox_normal(P.seed, P.counter ^ 0xAC710ull, (uint32_t)n, (uint32_t)d);
      a = a + P.dyn_rand * e * a;                                                           // :633
      a = clampf(a, -P.clip_actions, P.clip_actions);                                       // :636
      P.actions[i] = a;
    }
    if (P.torques) {
      const float2 pv = *reinterpret_cast<const float2*>(P.dof_state + 2 * i);
      const float lim = __ldg(P.torque_limits + d);
      const float t = __ldg(P.p_gains + i) * (a * P.action_scale + __ldg(P.default_dof_pos + d) - pv.x) - __ldg(P.d_gains + i) * pv.y;
      P.torques[i] = clampf(t, -lim, lim);                                                  // :921-925
    }
  }
}

}  // namespace hg

extern "C" int hg_action_torque(const HgActionArgs* a, void* stream) {
  using namespace hg;
  HG_CHECK(a && a->num_envs > 0 && a->D > 0 && a->actions, "hg_action_torque: bad arguments");
  HG_CHECK(!a->torques || (a->dof_state && a->p_gains && a->d_gains && a->default_dof_pos && a->torque_limits),
           "hg_action_torque: the torque law needs dof_state, gains, default_dof_pos and torque_limits");
  const long long total = (long long)a->num_envs * a->D;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  HG_CUDA(launch_kernel(action_torque_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, *a));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_env_step(const HgEnvConfig* cfg, const HgEnvBuffers* buf, void* stream) {
  using namespace hg;
  HG_CHECK(cfg && buf, "hg_env_step: null arguments");
  HG_CHECK(cfg->num_envs > 0 && cfg->D >= 12 && cfg->D <= 32, "hg_env_step: D must be in [12,32]");
  HG_CHECK(cfg->leg_offset >= 0 && cfg->leg_offset + 12 <= cfg->D, "hg_env_step: leg_offset + 12 must be <= D");
  HG_CHECK(11 + 3 * cfg->D <= MAX_FRAME && 25 + 4 * cfg->D <= MAX_FRAME, "hg_env_step: frame too wide");
  HG_CHECK(cfg->num_term_bodies >= 0 && cfg->num_term_bodies <= 4 && cfg->num_pen_bodies >= 0 && cfg->num_pen_bodies <= 4,
           "hg_env_step: at most 4 termination / penalised bodies");
  HG_CHECK(cfg->frame_stack >= 1 && cfg->c_frame_stack >= 1, "hg_env_step: frame stacks");
  HG_CHECK(cfg->num_rigid_bodies == 0 || (cfg->num_rigid_bodies >= 1 && cfg->rigid_feet[0] >= 0 && cfg->rigid_feet[0] < cfg->num_rigid_bodies &&
                                          cfg->rigid_feet[1] >= 0 && cfg->rigid_feet[1] < cfg->num_rigid_bodies &&
                                          cfg->rigid_knees[0] >= 0 && cfg->rigid_knees[0] < cfg->num_rigid_bodies &&
                                          cfg->rigid_knees[1] >= 0 && cfg->rigid_knees[1] < cfg->num_rigid_bodies),
           "hg_env_step: compact rigid-body table: indices outside [0, num_rigid_bodies)");
  HG_CHECK((cfg->phases & (HG_ENV_PHASE_REWARD | HG_ENV_PHASE_OBS)) != 0, "hg_env_step: phases must select REWARD and/or OBS");
  const bool ring = buf->ring_obs.live32 != nullptr;
  HG_CHECK(ring || (buf->obs_prev != buf->obs_next && buf->priv_prev != buf->priv_next), "hg_env_step: history in/out must not alias");
  HG_CHECK(buf->root_states && buf->dof_state && buf->contact_forces && buf->rigid_state && buf->commands && buf->torques &&
               buf->rand_push_force && buf->rand_push_torque && buf->env_frictions && buf->body_mass && buf->default_dof_pos &&
               buf->noise_scale_vec && buf->actions && buf->last_actions && buf->last_last_actions && buf->last_dof_vel &&
               buf->last_root_vel && buf->feet_air_time && buf->last_contacts && buf->feet_height && buf->last_feet_z &&
               buf->ref_dof_pos && buf->episode_length_buf && buf->episode_sums && buf->rew_buf && buf->reset_buf && buf->time_out_buf,
           "hg_env_step: null buffer");
  HG_CHECK(ring || (buf->obs_prev && buf->priv_prev && buf->obs_next && buf->priv_next), "hg_env_step: null history buffer");
  const int So_ = 11 + 3 * cfg->D, Sp_ = 25 + 4 * cfg->D;
  if (ring) {
    const HgHistoryRing* rr[2] = {&buf->ring_obs, &buf->ring_priv};
    const int width[2] = {So_, Sp_}, frames[2] = {cfg->frame_stack, cfg->c_frame_stack};
    for (int i = 0; i < 2; ++i) {
      const HgHistoryRing& R = *rr[i];
      HG_CHECK(R.live32 && R.live16 && R.arch32 && R.arch16, "hg_env_step: history ring %d: null buffer", i);
      HG_CHECK(R.width == width[i] && R.frames == frames[i] && R.pitch16 >= R.width && R.pitch16 % 8 == 0,
               "hg_env_step: history ring %d: width %d / frames %d / pitch16 %d do not match the env layout", i, R.width, R.frames, R.pitch16);
      HG_CHECK(buf->ring_row >= 1 && buf->ring_row + R.frames - 1 < R.slots, "hg_env_step: ring_row %d outside the ring (%d slots)", buf->ring_row, R.slots);
    }
    HG_CHECK(buf->ring_first_valid && buf->ring_hist_start, "hg_env_step: the history ring needs first_valid and hist_start");
  }
  HG_CHECK((buf->store_rewards != nullptr) == (buf->store_dones != nullptr) && (buf->store_rewards != nullptr) == (buf->store_values != nullptr),
           "hg_env_step: store_values / store_rewards / store_dones go together");
  HG_CHECK(!buf->store_rewards || (cfg->phases & HG_ENV_PHASE_REWARD), "hg_env_step: the transition store needs the reward phase");
  // staging table (see EnvStage): sources, floats per env, shared-memory offsets for E envs per block
  auto make_stage = [&](int E) {
    EnvStage L;
    memset(&L, 0, sizeof(L));
    const int D = cfg->D, NB = cfg->num_bodies, N = cfg->num_envs;
    int o = 0;
    auto take = [&](int n) { const int at = o; o += (n + 3) & ~3; return at; };
    auto seg = [&](int id, const void* src, int width) {
      L.src[id] = reinterpret_cast<const float*>(src);
      L.width[id] = width;
      L.dst[id] = take(E * width);
    };
    L.fobs = take(E * So_); L.fpriv = take(E * Sp_); L.terms = take(HG_NUM_REWARDS * E); L.s2 = take(E);
    L.reset = take(E); L.ep = take(E); L.tout = take(E); L.rg = take(E * 52); L.lc = take(E); L.nsv = take(So_); L.ddp = take(D);
    L.noise = take(E * ((So_ + 3) & ~3));
    seg(SEG_ROOT, buf->root_states, 13); seg(SEG_DOF, buf->dof_state, 2 * D); seg(SEG_CF, buf->contact_forces, NB * 3);
    seg(SEG_CMD, buf->commands, 4); seg(SEG_TQ, buf->torques, D);
    seg(SEG_ACT, buf->actions_in ? buf->actions_in : buf->actions, D);      // raw policy output (clipped at use) or applied actions
    seg(SEG_LA, buf->last_actions, D); seg(SEG_LLA, buf->last_last_actions, D); seg(SEG_LDV, buf->last_dof_vel, D);
    seg(SEG_REF, buf->ref_dof_pos, D); seg(SEG_LRV, buf->last_root_vel, 6); seg(SEG_PRE, buf->root_vel_prepush, 6);
    seg(SEG_PF, buf->rand_push_force, 3); seg(SEG_PT, buf->rand_push_torque, 3); seg(SEG_FRIC, buf->env_frictions, 1);
    seg(SEG_MASS, buf->body_mass, 1); seg(SEG_AIR, buf->feet_air_time, 2); seg(SEG_FH, buf->feet_height, 2);
    seg(SEG_LFZ, buf->last_feet_z, 2); seg(SEG_EPL, buf->episode_length_buf, 2 /* int64 as two words */);
    seg(SEG_HIST, ring ? buf->ring_hist_start : nullptr, 1); seg(SEG_SVAL, buf->store_values, 1);
    const bool rewards = (cfg->phases & HG_ENV_PHASE_REWARD) != 0;
    for (int k = 0; k < HG_NUM_REWARDS; ++k) seg(SEG_SUMS + k, rewards ? buf->episode_sums + (size_t)k * N : nullptr, 1);
    L.total = o;
    return L;
  };
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool configured = false;
  if (!configured) {       // 32 envs per block stage ~60 KB: above the 48 KB a kernel gets without asking
    HG_CUDA(cudaFuncSetAttribute(env_step_kernel<16, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    HG_CUDA(cudaFuncSetAttribute(env_step_kernel<16, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    HG_CUDA(cudaFuncSetAttribute(env_step_kernel<32, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  if (!ring) {
    const EnvStage L = make_stage(16);
    HG_CUDA(launch_kernel(env_step_kernel<16, 512, false>, dim3((cfg->num_envs + 15) / 16), dim3(512), (size_t)L.total * 4, st, 1, *cfg, *buf, L));
  } else if (cfg->num_envs <= 2 * 16 * num_sms()) {      // up to ~two blocks per SM: 16 envs per block (more SMs, shorter per-SM chains)
    const EnvStage L = make_stage(16);
    HG_CUDA(launch_kernel(env_step_kernel<16, 256, true>, dim3((cfg->num_envs + 15) / 16), dim3(256), (size_t)L.total * 4, st, 1, *cfg, *buf, L));
  } else {                                               // many envs: 32 per block (every lane of the role warps busy)
    const EnvStage L = make_stage(32);
    HG_CUDA(launch_kernel(env_step_kernel<32, 256, true>, dim3((cfg->num_envs + 31) / 32), dim3(256), (size_t)L.total * 4, st, 1, *cfg, *buf, L));
  }
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}
