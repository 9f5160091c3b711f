// Fused env-step tensor math: one launch replaces the several hundred ATen launches of
// XBotLFreeEnv.post_physics_step (reference humanoid/envs/custom/humanoid_env.py:780-806 and callees).
//
// One warp per env.  Lane i < D owns DOF i (joint-space sums are warp-shuffle reductions); per-env scalars
// are broadcast loads; the two new frames are assembled in shared memory and the warp then streams the
// history shift: obs_next[n, :] = clip([obs_prev[n, So:], new_frame]) with fully coalesced 128-byte
// accesses (15 x So floats written, 14 x So read per env; same for the 3-frame critic stack).
// Algorithmic HBM bytes per env-step (D=12): 4*(47*29 + 73*5) = 6,912 B history + ~0.9 KB state.
//
// Compiled with -fmad=false: every +,-,*,/ is a separately rounded IEEE op in the reference's order, so all
// threshold decisions (contacts, resets, dead-bands, feet clearance) are bit-identical to torch-CPU; only
// sin/cos/exp/atan2/asin and the order of the short DOF sums differ (<= a few ulp).
#include "hg_common.cuh"

namespace hg {

constexpr int ENV_WARPS = 8;
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

__device__ __forceinline__ float wrap_2pi(float a) {   // python-style a % (2*pi), then > pi -> -= 2*pi
  float m = fmodf(a, TWO_PI_F);
  if (m < 0.f) m += TWO_PI_F;
  if (m > PI_F) m -= TWO_PI_F;
  return m;
}

// isaacgym get_euler_xyz + the wrap of get_euler_xyz_tensor (humanoid_env.py:51-56)
__device__ __forceinline__ Vec3 euler_xyz_wrapped(const Quat& q) {
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

__device__ __forceinline__ float gait_sin(const HgEnvConfig& c, long long ep_len) {
  const float t = (float)ep_len * c.dt;
  const float phase = c.phase_mul_reciprocal ? t * (1.0f / c.cycle_time) : t / c.cycle_time;   // :692
  return sinf(TWO_PI_F * phase);
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

__global__ void __launch_bounds__(ENV_WARPS * 32) env_step_kernel(const __grid_constant__ HgEnvConfig C,
                                                                  const __grid_constant__ HgEnvBuffers B) {
  __shared__ float s_frames[ENV_WARPS][2 * MAX_FRAME];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * ENV_WARPS + warp;
  if (n >= C.num_envs) return;
  const int D = C.D, NB = C.num_bodies, off = C.leg_offset;
  const int So = 11 + 3 * D, Sp = 25 + 4 * D;
  float* f_obs = s_frames[warp];
  float* f_priv = s_frames[warp] + MAX_FRAME;
  const bool dl = lane < D;

  // ------------------------------------------------------------------ loads
  const float* root = B.root_states + (size_t)n * 13;
  const Quat q = {__ldg(root + 3), __ldg(root + 4), __ldg(root + 5), __ldg(root + 6)};
  const Vec3 lin = {__ldg(root + 7), __ldg(root + 8), __ldg(root + 9)};
  const Vec3 ang = {__ldg(root + 10), __ldg(root + 11), __ldg(root + 12)};
  const float root_z = __ldg(root + 2);
  const float cmd0 = __ldg(B.commands + n * 4), cmd1 = __ldg(B.commands + n * 4 + 1), cmd2 = __ldg(B.commands + n * 4 + 2);
  float dof_pos = 0.f, dof_vel = 0.f, act = 0.f, lact = 0.f, llact = 0.f, ldv = 0.f, tq = 0.f, ref_old = 0.f, dflt = 0.f;
  if (dl) {
    const float2 pv = *reinterpret_cast<const float2*>(B.dof_state + ((size_t)n * D + lane) * 2);
    dof_pos = pv.x;
    dof_vel = pv.y;
    act = B.actions[(size_t)n * D + lane];
    lact = B.last_actions[(size_t)n * D + lane];
    llact = B.last_last_actions[(size_t)n * D + lane];
    ldv = B.last_dof_vel[(size_t)n * D + lane];
    tq = __ldg(B.torques + (size_t)n * D + lane);
    ref_old = B.ref_dof_pos[(size_t)n * D + lane];
    dflt = __ldg(B.default_dof_pos + lane);
  }
  const float* cf = B.contact_forces + (size_t)n * NB * 3;
  const float* rg = B.rigid_state + (size_t)n * NB * 13;

  const bool doA = (C.phases & HG_ENV_PHASE_REWARD) != 0, doB = (C.phases & HG_ENV_PHASE_OBS) != 0;
  long long ep = B.episode_length_buf[n] + (doA ? 1 : 0);                       // :780

  // ------------------------------------------------------------------ base-frame quantities :784-788
  const Vec3 blv = quat_rotate_inverse(q, lin);
  const Vec3 bav = quat_rotate_inverse(q, ang);
  const Vec3 pg = quat_rotate_inverse(q, Vec3{0.f, 0.f, -1.f});
  const Vec3 eul = euler_xyz_wrapped(q);

  // ------------------------------------------------------------------ check_termination :811-816
  bool term = false, time_out = false, reset = false;
  if (doA) {
    for (int i = 0; i < C.num_term_bodies; ++i) {
      const float* f = cf + C.term_bodies[i] * 3;
      const float x = __ldg(f), y = __ldg(f + 1), z = __ldg(f + 2);
      term = term || (sqrtf(x * x + y * y + z * z) > 1.0f);
    }
    time_out = (float)ep > C.max_episode_length;
    reset = term || time_out;
  } else {
    reset = B.reset_buf[n] != 0;       // decided by an earlier HG_ENV_PHASE_REWARD launch
  }

  // ------------------------------------------------------------------ rewards :1170-1437
  const float s_pre = gait_sin(C, ep);
  float st0 = s_pre >= 0.f ? 1.f : 0.f, st1 = s_pre < 0.f ? 1.f : 0.f;
  if (fabsf(s_pre) < 0.1f) { st0 = 1.f; st1 = 1.f; }
  const float fz0 = __ldg(cf + C.feet[0] * 3 + 2), fz1 = __ldg(cf + C.feet[1] * 3 + 2);
  const bool ct0 = fz0 > 5.0f, ct1 = fz1 > 5.0f;
  float r[HG_NUM_REWARDS];
  float air0 = B.feet_air_time[n * 2], air1 = B.feet_air_time[n * 2 + 1];
  float fh0 = B.feet_height[n * 2], fh1 = B.feet_height[n * 2 + 1];
  const float foot_z0 = __ldg(rg + C.feet[0] * 13 + 2), foot_z1 = __ldg(rg + C.feet[1] * 13 + 2);
  const float fzz0 = foot_z0 - 0.05f, fzz1 = foot_z1 - 0.05f;
  float rew = 0.f;
  if (doA) {

  {  // action_smoothness :1427-1437
    const float d1 = lact - act, d2 = act + llact - 2.0f * lact;
    const float t1 = warp_sum(dl ? d1 * d1 : 0.f), t2 = warp_sum(dl ? d2 * d2 : 0.f), t3 = warp_sum(dl ? fabsf(act) : 0.f);
    r[HG_R_ACTION_SMOOTHNESS] = t1 + t2 + 0.05f * t3;
  }
  {  // base_acc :1283-1290
    const float* lrv = B.last_root_vel + (size_t)n * 6;
    const float a0 = lrv[0] - lin.x, a1 = lrv[1] - lin.y, a2 = lrv[2] - lin.z;
    const float a3 = lrv[3] - ang.x, a4 = lrv[4] - ang.y, a5 = lrv[5] - ang.z;
    r[HG_R_BASE_ACC] = expf(-sqrtf(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5) * 3.0f);
  }
  {  // base_height :1271-1281
    const float meas = (foot_z0 * st0 + foot_z1 * st1) / (st0 + st1);
    const float bh = root_z - (meas - 0.05f);
    r[HG_R_BASE_HEIGHT] = expf(-fabsf(bh - C.base_height_target) * 100.0f);
  }
  {  // collision :1419-1425
    float cnt = 0.f;
    for (int i = 0; i < C.num_pen_bodies; ++i) {
      const float* f = cf + C.pen_bodies[i] * 3;
      const float x = __ldg(f), y = __ldg(f + 1), z = __ldg(f + 2);
      cnt += (sqrtf(x * x + y * y + z * z) > 0.1f) ? 1.f : 0.f;
    }
    r[HG_R_COLLISION] = cnt;
  }
  const float jd = dof_pos - dflt;
  {  // default_joint_pos :1259-1269
    const float jd2 = dl ? jd * jd : 0.f;
    const float l0 = __shfl_sync(0xffffffffu, jd2, off), l1 = __shfl_sync(0xffffffffu, jd2, off + 1);
    const float r0 = __shfl_sync(0xffffffffu, jd2, off + 6), r1 = __shfl_sync(0xffffffffu, jd2, off + 7);
    float yr = sqrtf(l0 + l1) + sqrtf(r0 + r1);
    yr = clampf(yr - 0.1f, 0.f, 50.f);
    r[HG_R_DEFAULT_JOINT_POS] = expf(-yr * 100.0f) - 0.01f * sqrtf(warp_sum(jd2));
  }
  {  // dof_acc :1412-1417, dof_vel :1405-1410, torques :1398-1403
    const float acc = (ldv - dof_vel) / C.dt;
    r[HG_R_DOF_ACC] = warp_sum(dl ? acc * acc : 0.f);
    r[HG_R_DOF_VEL] = warp_sum(dl ? dof_vel * dof_vel : 0.f);
    r[HG_R_TORQUES] = warp_sum(dl ? tq * tq : 0.f);
  }
  {  // feet_air_time :1216-1230 (stateful)
    const bool lc0 = B.last_contacts[n * 2] != 0, lc1 = B.last_contacts[n * 2 + 1] != 0;
    const bool cfl0 = ct0 || (st0 != 0.f) || lc0, cfl1 = ct1 || (st1 != 0.f) || lc1;
    const float first0 = (air0 > 0.f && cfl0) ? 1.f : 0.f, first1 = (air1 > 0.f && cfl1) ? 1.f : 0.f;
    air0 += C.dt;
    air1 += C.dt;
    r[HG_R_FEET_AIR_TIME] = clampf(air0, 0.f, 0.5f) * first0 + clampf(air1, 0.f, 0.5f) * first1;
    air0 *= cfl0 ? 0.f : 1.f;
    air1 *= cfl1 ? 0.f : 1.f;
  }
  {  // feet_clearance :1342-1363 (stateful)
    fh0 += fzz0 - B.last_feet_z[n * 2];
    fh1 += fzz1 - B.last_feet_z[n * 2 + 1];
    const float p0 = fabsf(fh0 - C.target_feet_height) < 0.01f ? 1.f : 0.f;
    const float p1 = fabsf(fh1 - C.target_feet_height) < 0.01f ? 1.f : 0.f;
    r[HG_R_FEET_CLEARANCE] = p0 * (1.f - st0) + p1 * (1.f - st1);
    fh0 *= ct0 ? 0.f : 1.f;
    fh1 *= ct1 ? 0.f : 1.f;
  }
  {  // feet_contact_forces :1251-1257, feet_contact_number :1232-1240
    const float* f0 = cf + C.feet[0] * 3;
    const float* f1 = cf + C.feet[1] * 3;
    const float n0 = sqrtf(__ldg(f0) * __ldg(f0) + __ldg(f0 + 1) * __ldg(f0 + 1) + fz0 * fz0);
    const float n1 = sqrtf(__ldg(f1) * __ldg(f1) + __ldg(f1 + 1) * __ldg(f1 + 1) + fz1 * fz1);
    r[HG_R_FEET_CONTACT_FORCES] = clampf(n0 - C.max_contact_force, 0.f, 400.f) + clampf(n1 - C.max_contact_force, 0.f, 400.f);
    const float e0 = ((ct0 ? 1.f : 0.f) == st0) ? 1.f : -0.3f, e1 = ((ct1 ? 1.f : 0.f) == st1) ? 1.f : -0.3f;
    r[HG_R_FEET_CONTACT_NUMBER] = (e0 + e1) / 2.0f;
  }
  {  // feet_distance :1180-1190, knee_distance :1192-1202
    auto dist_reward = [&](int b0, int b1, float max_d) {
      const float dx = __ldg(rg + b0 * 13) - __ldg(rg + b1 * 13), dy = __ldg(rg + b0 * 13 + 1) - __ldg(rg + b1 * 13 + 1);
      const float d = sqrtf(dx * dx + dy * dy);
      const float d_min = clampf(d - C.min_dist, -0.5f, 0.f), d_max = clampf(d - max_d, 0.f, 0.5f);
      return (expf(-fabsf(d_min) * 100.0f) + expf(-fabsf(d_max) * 100.0f)) / 2.0f;
    };
    r[HG_R_FEET_DISTANCE] = dist_reward(C.feet[0], C.feet[1], C.max_dist);
    r[HG_R_KNEE_DISTANCE] = dist_reward(C.knees[0], C.knees[1], C.max_dist / 2.0f);
  }
  {  // foot_slip :1204-1214 (columns 10:12 as written in the reference)
    const float ax = __ldg(rg + C.feet[0] * 13 + 10), ay = __ldg(rg + C.feet[0] * 13 + 11);
    const float bx = __ldg(rg + C.feet[1] * 13 + 10), by = __ldg(rg + C.feet[1] * 13 + 11);
    r[HG_R_FOOT_SLIP] = sqrtf(sqrtf(ax * ax + ay * ay)) * (ct0 ? 1.f : 0.f) + sqrtf(sqrtf(bx * bx + by * by)) * (ct1 ? 1.f : 0.f);
  }
  {  // joint_pos :1170-1178, with the PREVIOUS step's reference trajectory
    const float d = dof_pos - ref_old;
    const float nrm = sqrtf(warp_sum(dl ? d * d : 0.f));
    r[HG_R_JOINT_POS] = expf(-2.0f * nrm) - 0.2f * clampf(nrm, 0.f, 0.5f);
  }
  {  // low_speed :1365-1396
    const float av = fabsf(blv.x), ac = fabsf(cmd0);
    const bool low = av < 0.5f * ac, high = av > 1.2f * ac;
    float v = 0.f;
    if (low) v = -1.0f;
    if (high) v = 0.f;
    if (!(low || high)) v = 1.2f;
    if (sgnf(blv.x) != sgnf(cmd0)) v = -2.0f;
    r[HG_R_LOW_SPEED] = v * (ac > 0.1f ? 1.f : 0.f);
  }
  r[HG_R_ORIENTATION] = (expf(-(fabsf(eul.x) + fabsf(eul.y)) * 10.0f) + expf(-sqrtf(pg.x * pg.x + pg.y * pg.y) * 20.0f)) / 2.0f;
  {  // track_vel_hard :1304-1321, tracking_* :1323-1340, vel_mismatch_exp :1292-1302
    const float ex = cmd0 - blv.x, ey = cmd1 - blv.y;
    const float e_l = sqrtf(ex * ex + ey * ey), e_a = fabsf(cmd2 - bav.z);
    r[HG_R_TRACK_VEL_HARD] = (expf(-e_l * 10.0f) + expf(-e_a * 10.0f)) / 2.0f - 0.2f * (e_l + e_a);
    const float ea2 = (cmd2 - bav.z) * (cmd2 - bav.z);
    r[HG_R_TRACKING_ANG_VEL] = expf(-ea2 * C.tracking_sigma);
    r[HG_R_TRACKING_LIN_VEL] = expf(-(ex * ex + ey * ey) * C.tracking_sigma);
    r[HG_R_VEL_MISMATCH_EXP] = (expf(-(blv.z * blv.z) * 10.0f) + expf(-sqrtf(bav.x * bav.x + bav.y * bav.y) * 5.0f)) / 2.0f;
  }
  // total in alphabetical order (:894-898), episode sums, clip (:899-900)
#pragma unroll
  for (int k = 0; k < HG_NUM_REWARDS; ++k) {
    const float sc = C.reward_scale_dt[k];
    if (sc != 0.f) {
      const float term_k = r[k] * sc;
      rew += term_k;
      if (lane == (k & 31)) {
        const float prev = B.episode_sums[(size_t)k * C.num_envs + n];
        B.episode_sums[(size_t)k * C.num_envs + n] = (reset && doB) ? 0.f : prev + term_k;   // zeroed by reset_idx (:1144)
        if (B.reward_terms) B.reward_terms[(size_t)k * C.num_envs + n] = term_k;
      }
    }
  }
  if (C.only_positive_rewards) rew = fmaxf(rew, 0.f);
  if (lane == 0) {
    B.last_contacts[n * 2] = ct0; B.last_contacts[n * 2 + 1] = ct1;
    B.feet_height[n * 2] = fh0; B.feet_height[n * 2 + 1] = fh1;
    B.last_feet_z[n * 2] = fzz0; B.last_feet_z[n * 2 + 1] = fzz1;
    B.rew_buf[n] = rew;
    B.reset_buf[n] = reset ? 1 : 0;
    B.time_out_buf[n] = time_out ? 1 : 0;
  }
  }  // doA
  if (!doB) {                          // reward-only launch: leave reset handling and observations to the next launch
    if (lane == 0) {
      B.feet_air_time[n * 2] = air0; B.feet_air_time[n * 2 + 1] = air1;
      B.episode_length_buf[n] = ep;
    }
    return;
  }
  if (!doA && reset && lane < HG_NUM_REWARDS) B.episode_sums[(size_t)lane * C.num_envs + n] = 0.f;

  // ------------------------------------------------------------------ reset_idx buffer part :1132-1139
  if (reset) {
    act = 0.f; lact = 0.f; llact = 0.f; ldv = 0.f;
    air0 = 0.f; air1 = 0.f;
    ep = 0;
  }

  // ------------------------------------------------------------------ compute_observations :818-887
  const float ph = gait_phase(C, ep);
  const float s2 = sinf(TWO_PI_F * ph), c2 = cosf(TWO_PI_F * ph);
  float so0 = s2 >= 0.f ? 1.f : 0.f, so1 = s2 < 0.f ? 1.f : 0.f;
  const bool dsup = fabsf(s2) < 0.1f;
  if (dsup) { so0 = 1.f; so1 = 1.f; }
  // compute_ref_state :717-744
  float ref_new = 0.f;
  {
    const float sl = s2 > 0.f ? 0.f : s2, sr = s2 < 0.f ? 0.f : s2;
    const float sc1 = C.target_joint_pos_scale, sc2 = 2.0f * C.target_joint_pos_scale;
    const int j = lane - off;
    if (j == 2 || j == 4) ref_new = sl * sc1;
    if (j == 3) ref_new = sl * sc2;
    if (j == 8 || j == 10) ref_new = sr * sc1;
    if (j == 9) ref_new = sr * sc2;
    if (dsup || !dl) ref_new = 0.f;
  }
  if (lane == 0) {
    f_obs[0] = s2; f_obs[1] = c2;
    f_obs[2] = cmd0 * C.lin_vel_scale; f_obs[3] = cmd1 * C.lin_vel_scale; f_obs[4] = cmd2 * C.ang_vel_scale;
    f_obs[5 + 3 * D + 0] = bav.x * C.ang_vel_scale; f_obs[5 + 3 * D + 1] = bav.y * C.ang_vel_scale; f_obs[5 + 3 * D + 2] = bav.z * C.ang_vel_scale;
    f_obs[8 + 3 * D + 0] = eul.x * C.quat_scale; f_obs[8 + 3 * D + 1] = eul.y * C.quat_scale; f_obs[8 + 3 * D + 2] = eul.z * C.quat_scale;
  }
  if (lane == 1) {
    float* p = f_priv;
    p[0] = s2; p[1] = c2;
    p[2] = cmd0 * C.lin_vel_scale; p[3] = cmd1 * C.lin_vel_scale; p[4] = cmd2 * C.ang_vel_scale;
    float* t = p + 5 + 4 * D;
    t[0] = blv.x * C.lin_vel_scale; t[1] = blv.y * C.lin_vel_scale; t[2] = blv.z * C.lin_vel_scale;
    t[3] = bav.x * C.ang_vel_scale; t[4] = bav.y * C.ang_vel_scale; t[5] = bav.z * C.ang_vel_scale;
    t[6] = eul.x * C.quat_scale; t[7] = eul.y * C.quat_scale; t[8] = eul.z * C.quat_scale;
    t[9] = __ldg(B.rand_push_force + n * 3); t[10] = __ldg(B.rand_push_force + n * 3 + 1);
    t[11] = __ldg(B.rand_push_torque + n * 3); t[12] = __ldg(B.rand_push_torque + n * 3 + 1); t[13] = __ldg(B.rand_push_torque + n * 3 + 2);
    t[14] = __ldg(B.env_frictions + n);
    t[15] = __ldg(B.body_mass + n) / 30.0f;
    t[16] = so0; t[17] = so1;
    t[18] = ct0 ? 1.f : 0.f; t[19] = ct1 ? 1.f : 0.f;
  }
  if (dl) {
    const float qv = (dof_pos - dflt) * C.dof_pos_scale, dqv = dof_vel * C.dof_vel_scale;
    f_obs[5 + lane] = qv; f_obs[5 + D + lane] = dqv; f_obs[5 + 2 * D + lane] = act;
    f_priv[5 + lane] = qv; f_priv[5 + D + lane] = dqv; f_priv[5 + 2 * D + lane] = act;
    f_priv[5 + 3 * D + lane] = dof_pos - ref_new;
  }
  __syncwarp();

  // ------------------------------------------------------------------ history shift + new frame + clip (:883-887, :654-657)
  {
    const int keep = (C.frame_stack - 1) * So, tot = C.frame_stack * So;
    const float* prev = B.obs_prev + (size_t)n * B.obs_prev_ld + So;
    float* next = B.obs_next + (size_t)n * B.obs_next_ld;
    for (int c = lane; c < tot; c += 32) {
      float v;
      if (c < keep) {
        v = reset ? 0.f : prev[c];
      } else {
        const int j = c - keep;
        v = f_obs[j];
        const float nv = __ldg(B.noise_scale_vec + j);
        if (C.noise_level != 0.f) {
          const float e = B.noise ? __ldg(B.noise + (size_t)n * So + j) : (nv != 0.f ? env_philox_normal(B.seed, B.counter, n, j) : 0.f);
          v = v + e * nv * C.noise_level;                                          // :877
        }
      }
      next[c] = clampf(v, -C.clip_obs, C.clip_obs);
    }
  }
  {
    const int keep = (C.c_frame_stack - 1) * Sp, tot = C.c_frame_stack * Sp;
    const float* prev = B.priv_prev + (size_t)n * B.priv_prev_ld + Sp;
    float* next = B.priv_next + (size_t)n * B.priv_next_ld;
    for (int c = lane; c < tot; c += 32) {
      const float v = c < keep ? (reset ? 0.f : prev[c]) : f_priv[c - keep];
      next[c] = clampf(v, -C.clip_obs, C.clip_obs);
    }
  }

  // ------------------------------------------------------------------ state hand-over :802-806 and outputs
  if (dl) {
    const size_t i = (size_t)n * D + lane;
    B.last_last_actions[i] = lact;
    B.last_actions[i] = act;
    B.actions[i] = act;
    B.last_dof_vel[i] = dof_vel;
    B.ref_dof_pos[i] = ref_new;
  }
  if (lane < 6) B.last_root_vel[(size_t)n * 6 + lane] = __ldg(root + 7 + lane);
  if (lane == 0) {
    B.feet_air_time[n * 2] = air0; B.feet_air_time[n * 2 + 1] = air1;
    B.episode_length_buf[n] = ep;
  }
}

}  // namespace hg

extern "C" int hg_env_step(const HgEnvConfig* cfg, const HgEnvBuffers* buf, void* stream) {
  using namespace hg;
  HG_CHECK(cfg && buf, "hg_env_step: null arguments");
  HG_CHECK(cfg->num_envs > 0 && cfg->D >= 12 && cfg->D <= 32, "hg_env_step: D must be in [12,32]");
  HG_CHECK(cfg->leg_offset >= 0 && cfg->leg_offset + 12 <= cfg->D, "hg_env_step: leg_offset + 12 must be <= D");
  HG_CHECK(11 + 3 * cfg->D <= MAX_FRAME && 25 + 4 * cfg->D <= MAX_FRAME, "hg_env_step: frame too wide");
  HG_CHECK(cfg->num_term_bodies >= 0 && cfg->num_term_bodies <= 4 && cfg->num_pen_bodies >= 0 && cfg->num_pen_bodies <= 4,
           "hg_env_step: at most 4 termination / penalised bodies");
  HG_CHECK(cfg->frame_stack >= 1 && cfg->c_frame_stack >= 1, "hg_env_step: frame stacks");
  HG_CHECK((cfg->phases & (HG_ENV_PHASE_REWARD | HG_ENV_PHASE_OBS)) != 0, "hg_env_step: phases must select REWARD and/or OBS");
  HG_CHECK(buf->obs_prev != buf->obs_next && buf->priv_prev != buf->priv_next, "hg_env_step: history in/out must not alias");
  HG_CHECK(buf->root_states && buf->dof_state && buf->contact_forces && buf->rigid_state && buf->commands && buf->torques &&
               buf->rand_push_force && buf->rand_push_torque && buf->env_frictions && buf->body_mass && buf->default_dof_pos &&
               buf->noise_scale_vec && buf->actions && buf->last_actions && buf->last_last_actions && buf->last_dof_vel &&
               buf->last_root_vel && buf->feet_air_time && buf->last_contacts && buf->feet_height && buf->last_feet_z &&
               buf->ref_dof_pos && buf->episode_length_buf && buf->episode_sums && buf->obs_prev && buf->priv_prev &&
               buf->obs_next && buf->priv_next && buf->rew_buf && buf->reset_buf && buf->time_out_buf,
           "hg_env_step: null buffer");
  const int blocks = (cfg->num_envs + ENV_WARPS - 1) / ENV_WARPS;
  env_step_kernel<<<blocks, ENV_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*cfg, *buf);
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}
