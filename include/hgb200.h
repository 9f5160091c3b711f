/* hgb200.h -- C ABI of libhgb200.so: the B200 (sm_100a) kernels of humanoid-gym's PPO hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer unless the
 * name ends in _host.  `stream` is a cudaStream_t passed as void*.  Every function returns
 * HG_OK (0) or a negative HG_ERR_* code; hg_last_error() returns the text.  There is no CPU
 * fallback anywhere: a call that cannot run on the GPU fails.
 *
 * Each entry point names the reference code it replaces (paths relative to the reference root,
 * Rengar-Yang/humanoid-gym-with-comments).
 */
#ifndef HGB200_H_
#define HGB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_OK 0
#define HG_ERR_INVALID (-1)
#define HG_ERR_CUDA (-2)

#define HG_ABI_VERSION 8

const char* hg_last_error(void);
int hg_abi_version(void);
/* Number of kernel launches issued through this library since load (bench.py's gpu_launches). */
int64_t hg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense contractions: grouped bf16 GEMM on tcgen05 tensor cores (TMA loads, TMEM accumulators).
 * Replaces every nn.Linear forward/backward of the three MLPs (humanoid/algo/ppo/actor_critic.py:
 * 62,76,89 via torch autograd) and the ELU forward/backward between them (:44).
 *
 *   C[M,N] = A[M,K] * B[N,K]^T      fp32 accumulate
 *
 * A is row-major [M,lda] ("K-major") or, with a_mn_major=1, stored as [K,lda] with M contiguous.
 * Same for B with N.  Leading dimensions are in elements and must be multiples of 8 (16 bytes).
 * Up to HG_MAX_GROUP independent problems run in ONE launch (actor / critic / lin-vel layers).
 */
#define HG_MAX_GROUP 4

enum HgEpilogue {
  HG_EPI_STORE_F32 = 0,     /* out f32 [M,ldo] = acc (+ bias[n] if bias)                              */
  HG_EPI_BIAS_ELU_BF16 = 1, /* out bf16 [M,ldo] = elu(acc + bias[n])        (Linear+ELU forward)      */
  HG_EPI_DELU_BF16 = 2,     /* out bf16 [M,ldo] = acc * elu'(act[m,n]); colsum[n] += column sums      */
                            /*   (dX of a Linear, times ELU' of the layer below, + its bias gradient) */
  HG_EPI_ATOMIC_F32 = 3,    /* out f32 [M,ldo] += acc   (split-K weight gradient)                     */
  HG_EPI_BIAS_BF16 = 4,     /* out bf16 [M,ldo] = acc + bias[n]                                       */
  HG_EPI_BIAS_ELU_HEAD = 5  /* elu(acc + bias[n]) is NOT stored: it feeds a fused output layer (HgFusedHead)  */
};

typedef struct HgGemmProblem {
  const void* A;   /* bf16 */
  const void* B;   /* bf16 */
  int64_t lda, ldb;
  int32_t a_mn_major, b_mn_major;
  int32_t M, N, K;
  int32_t k_splits; /* >=1; >1 only with HG_EPI_ATOMIC_F32 */
  void* out;
  int64_t ldo;
  const float* bias;  /* [N] or NULL */
  const void* act;    /* bf16 [M,ldact] (HG_EPI_DELU_BF16) */
  int64_t ldact;
  float* colsum;      /* [N] or NULL (HG_EPI_DELU_BF16) */
  /* HG_EPI_ATOMIC_F32 only, 0 = off: the N columns are frames of out_frame_src columns of which the first out_frame_dst
   * are real (the rest is padding): column f*out_frame_src + c goes to out[., f*out_frame_dst + c], c >= out_frame_dst is
   * dropped.  (Weight gradient of a first layer whose input rows come from the frame-pitched history ring.) */
  int32_t out_frame_src, out_frame_dst;
} HgGemmProblem;

/* block_n: 128 or 256 (UMMA N of one accumulator tile).
 * cta_group: 1 = one CTA per 128 x block_n tile; 2 = a CTA pair per 256 x block_n tile (tcgen05.mma.cta_group::2:
 * each CTA loads half of B, the pair shares it through distributed shared memory). */
int hg_gemm_grouped(const HgGemmProblem* problems_host, int num_problems, int epilogue, int block_n, int cta_group,
                    void* stream);

/* The same launch with the host-side preparation (argument checks, TMA tensor-map encoding, tile schedule) done
 * once: the rollout issues the identical launches for every storage slot on every iteration, so the host
 * keeps one plan per (slot, depth) and replays it with a single cheap call. */
typedef struct HgGemmPlan HgGemmPlan;
int hg_gemm_plan_create(const HgGemmProblem* problems_host, int num_problems, int epilogue, int block_n, int cta_group,
                        HgGemmPlan** out);
int hg_gemm_plan_run(const HgGemmPlan* plan, void* stream);
void hg_gemm_plan_destroy(HgGemmPlan* plan);

/* Rollout only: the LAST hidden layer of each net together with its output layer and everything PPO.act does with it
 * (actor_critic.py:127-148, ppo.py:116-126).  One 128-row tile holds complete rows of the last hidden layer (N <= 128),
 * so the epilogue rounds elu(acc + bias) to bf16 (the value the update's forward pass will see), multiplies by the
 * fp32 output weights and
 *   HG_HEAD_POLICY: mu = W h + b, a = mu + std * eps, log-prob of a        -> out (mu), actions, sigma, log_prob
 *   HG_HEAD_LINEAR: out = W h + b                                         (critic value)
 * eps is an optional [M, n_out] tensor of externally drawn normals; NULL -> Philox4x32-10 keyed by (seed, counter),
 * the same stream as hg_policy_head_rollout.  Requires block_n = 128, cta_group = 1, N <= 128, N % 32 == 0, n_out <= 16. */
enum HgHeadKind { HG_HEAD_NONE = 0, HG_HEAD_POLICY = 1, HG_HEAD_LINEAR = 2 };
typedef struct HgFusedHead {
  int32_t kind, n_out;
  const float* W;      /* f32 [n_out, N] */
  const float* b;      /* f32 [n_out] */
  float* out;          /* f32 [M, n_out] */
  const float* std;    /* POLICY: f32 [n_out] */
  float* actions;      /* POLICY: f32 [M, n_out] */
  float* sigma;        /* POLICY: f32 [M, n_out] */
  float* log_prob;     /* POLICY: f32 [M] */
} HgFusedHead;
int hg_gemm_plan_create_heads(const HgGemmProblem* problems_host, const HgFusedHead* heads_host, int num_problems,
                              HgGemmPlan** out);
int hg_gemm_plan_run_heads(const HgGemmPlan* plan, const float* eps, uint64_t seed, uint64_t counter, void* stream);
/* The whole `PPO.act` forward of one env step (ppo.py:116-126) in one call: hg_gemm_plan_run(plans[i]) for i < n, then
 * hg_gemm_plan_run_heads(head_plan, ...).  Identical launches; one binding call instead of n + 1. */
int hg_rollout_forward(const HgGemmPlan* const* plans, int n, const HgGemmPlan* head_plan, const float* eps,
                       uint64_t seed, uint64_t counter, void* stream);

/* Rollout, nets with three hidden layers: layer 2, layer 3 and the output layer of up to two nets (actor, critic) in ONE
 * launch -- what hg_gemm_plan_create (layer 2) + hg_gemm_plan_create_heads (layer 3 + head) do in two.  A CTA keeps its
 * 128-row tile of h2 = elu(h1 W2^T + b2) in shared memory as the bf16 K-major operand of layer 3, so h2 is never stored.
 *   layer2[g]: A = h1 [M,K1] bf16, B = W2 [N2,K1] bf16, bias = b2, M, N = N2 (64..256, multiple of 64), K = K1
 *   layer3[g]: B = W3 [N3,N2] bf16, bias = b3, N = N3 (32..128, multiple of 32), K = N2      (A, M, out unused)
 * The output layer runs on the tensor core too (W_out = hi + mid + lo bf16 terms, exact products, fp32 accumulate).
 * Run it with hg_gemm_plan_run_heads; results match the two-launch path to fp32 round-off (2e-6). */
int hg_gemm_plan_create_tail(const HgGemmProblem* layer2_host, const HgGemmProblem* layer3_host, const HgFusedHead* heads_host,
                             int num_nets, HgGemmPlan** out);

/* The whole rollout forward of one env step -- layer 1 AND the tail above -- in ONE launch (`rollout_step_kernel`): the first
 * CTAs run the grouped layer-1 GEMM (layer1[g]: obs / critic-obs window -> h1 of net g, bias + ELU, bf16), the others each own
 * one (128-row block, net) of the tail and start once the layer-1 tiles of their row block are published (per-tile counters
 * in device memory owned by the plan; the layer-1 CTAs never wait).  Requirements: those of the tail; layer1[g].out must be
 * layer2[g].A (same M, pitch), N a multiple of 32, and the tail CTAs must leave >= 32 SMs for layer 1 (otherwise the call
 * fails and the caller keeps the two launches).  Same results as the two launches, bit for bit.  Run with
 * hg_gemm_plan_run_heads / hg_rollout_forward.  Measured on B200 at 4096 rows: SLOWER than the two launches (32.5 vs 26 us per
 * step: layer 1 is per-SM fill-bound and loses the tail's 64 SMs), so the Python mirror uses it only with HG_ROLLOUT_STEP=1. */
int hg_gemm_plan_create_step(const HgGemmProblem* layer1_host, const HgGemmProblem* layer2_host, const HgGemmProblem* layer3_host,
                             const HgFusedHead* heads_host, int num_nets, HgGemmPlan** out);

/* ------------------------------------------------------------------------------------------
 * GAE: replaces RolloutStorage.compute_returns (humanoid/algo/ppo/rollout_storage.py:122-143).
 * rewards/values/returns/advantages are [T,N] fp32, dones [T,N] uint8, last_values [N].
 * One reverse-time scan per env, then mean / unbiased std over all T*N and normalisation.
 * scratch: 4 doubles (zeroed by the call).  Bit-exact with the reference's fp32 op order for
 * `returns`; advantages agree to fp32 rounding of the mean/std reduction.
 */
int hg_gae_returns(const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
                   float gamma, float lam, int T, int N, float* returns, float* advantages, double* scratch,
                   void* stream);

/* ------------------------------------------------------------------------------------------
 * Rollout-side learner glue.
 */
/* obs ingest: replaces the two big copy_ of RolloutStorage.add_transitions (rollout_storage.py:90-91)
 * and produces the bf16, 16-byte-row-aligned shadow the tensor-core layers read.
 * src f32 [rows, src_ld] -> dst_f32 [rows, dst_ld] (may be NULL) and dst_bf16 [rows, ld16] with the
 * columns [cols, ld16) zero-filled. */
int hg_ingest_rows(const float* src, int64_t src_ld, int rows, int cols, float* dst_f32, int64_t dst_ld,
                   void* dst_bf16, int64_t ld16, void* stream);

/* Minibatch gather (rollout_storage.py:176-181): dst[r,:] = src[idx[r],:] for bf16 rows of `ld` elements. */
int hg_gather_rows_bf16(const void* src, int64_t ld, const int64_t* idx, int rows, void* dst, void* stream);

/* Generic `ActorCritic.act` / `get_actions_log_prob` (humanoid/algo/ppo/actor_critic.py:127-138) outside the fused rollout path:
 *   hg_gaussian_sample:   actions = mu + std * eps  (eps [n,A] given, or NULL -> Philox4x32-10 keyed by (seed, counter, row, action)),
 *                         log_prob (optional) = sum_a log N(actions | mu, std)           (Normal.sample / Normal.log_prob)
 *   hg_gaussian_log_prob: out[n] = sum_a log N(actions | mu, std)
 * mu, actions: f32 [n, A]; std: f32 [A]. */
int hg_gaussian_sample(const float* mu, const float* std, const float* eps, uint64_t seed, uint64_t counter, int n, int A,
                       float* actions, float* log_prob, void* stream);
int hg_gaussian_log_prob(const float* actions, const float* mu, const float* std, int n, int A, float* out, void* stream);

/* Output layers of actor and critic + Gaussian sampling + log-prob for one rollout step.
 * Replaces ActorCritic.act / evaluate / get_actions_log_prob as used by PPO.act
 * (actor_critic.py:127-138,144-146; ppo.py:116-126) for the last Linear of each net.
 *   h_actor  bf16 [n, ld_ha] (last hidden activations, width H_a), W_a f32 [A,H_a], b_a [A], std [A]
 *   h_critic bf16 [n, ld_hc], W_c f32 [1,H_c], b_c [1]
 *   eps f32 [n,A] ~ N(0,1) or NULL (then Philox4x32-10 keyed by (seed, counter, row) is used)
 * outputs (all f32): actions [n,A] = mu + std*eps, mu [n,A], sigma [n,A], log_prob [n], values [n]. */
int hg_policy_head_rollout(const void* h_actor, int64_t ld_ha, int H_a, const float* W_a, const float* b_a,
                           const float* std, const void* h_critic, int64_t ld_hc, int H_c, const float* W_c,
                           const float* b_c, const float* eps, uint64_t seed, uint64_t counter, int n, int A,
                           float* actions, float* mu, float* sigma, float* log_prob, float* values, void* stream);

/* Value-only head (PPO.compute_returns' critic pass on the last observation, ppo.py:140-141). */
int hg_value_head(const void* h_critic, int64_t ld_hc, int H_c, const float* W_c, const float* b_c, int n,
                  float* values, void* stream);

/* Generic fp32 output layer: out[n, n_out] = h[n,:H] * W[n_out,H]^T + b (ActorCritic.act_inference / evaluate /
 * base_get_lin_vel, actor_critic.py:140-148, for callers outside the fused PPO paths). */
int hg_head_forward(const void* h, int64_t ld_h, int H, const float* W, const float* b, int n_out, int n,
                    float* out, void* stream);

/* rewards + gamma * values * time_outs and the uint8 dones (ppo.py:129-133, rollout_storage.py:93-94). */
int hg_store_rewards_dones(const float* rewards, const float* values, const uint8_t* time_outs,
                           const uint8_t* dones, float gamma, int n, float* out_rewards, uint8_t* out_dones,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Update-side: output layers + all PPO losses + their gradients, one launch per minibatch.
 * Replaces ppo.py:154-208 (log-prob, KL, ratio, clipped surrogate, clipped value loss, entropy bonus,
 * lin-vel MSE) and the autograd backward of the three output Linear layers and of the ELU below them.
 */
typedef struct HgHeadNet {
  const void* h;     /* bf16 [rows, ld_h]: last hidden activations of this net for the minibatch */
  int64_t ld_h;
  int32_t H;         /* hidden width (<=256, multiple of 32) */
  int32_t n_out;     /* 12 / 1 / 3 */
  const float* W;    /* f32 [n_out, H] */
  const float* b;    /* f32 [n_out] */
  float* gW;         /* f32 [n_out, H], accumulated atomically */
  float* gb;         /* f32 [n_out] */
  void* dZ;          /* bf16 [rows, ld_dz]: dLoss/d(pre-activation of the last hidden layer) */
  int64_t ld_dz;
  float* gb_hidden;  /* f32 [H]: bias gradient of the last hidden layer (column sums of dZ) */
} HgHeadNet;

typedef struct HgLossArgs {
  HgHeadNet actor, critic, linvel;
  int32_t rows;            /* minibatch size */
  int32_t A;               /* num actions */
  const int64_t* idx;      /* [rows] flat row indices t*N+n into the storage tensors below */
  const float* actions;    /* [TN, A] */
  const float* old_mu;     /* [TN, A] */
  const float* old_sigma;  /* [TN, A] */
  const float* old_log_prob; /* [TN] */
  const float* advantages; /* [TN] */
  const float* returns;    /* [TN] */
  const float* old_values; /* [TN] */
  const float* critic_obs; /* f32 [TN, ld_cobs] : lin-vel target = cols [lv_col, lv_col+3) (rollout_storage.py:164) */
  int64_t ld_cobs;
  int32_t lv_col;
  const float* std;        /* [A] */
  float* g_std;            /* [A] */
  float clip_param, value_loss_coef, entropy_coef, lin_vel_coef;
  int32_t use_clipped_value_loss;
  float inv_world;         /* 1/world_size: gradients are pre-scaled so that a SUM all-reduce yields the global mean */
  /* stats[0..4] += {sum value-loss terms, sum surrogate terms, sum lin-vel sq err, sum kl, sum entropy} over the
   * minibatch rows (fp32; a few atomics per thread block).  Lives in the 8 floats that FOLLOW the gradient arena so
   * that the single gradient all-reduce also reduces the KL (every rank derives the same learning rate). */
  float* stats;
} HgLossArgs;

int hg_ppo_head_loss(const HgLossArgs* args_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser: replaces nn.utils.clip_grad_norm_ + torch.optim.Adam.step + the adaptive-KL learning
 * rate (ppo.py:162-174, 211-214) on flat fp32 arenas, with no host synchronisation.
 *   grads: [count + HG_STATS_SLOTS] floats; grads[count + j] are the loss statistics of hg_ppo_head_loss.
 *   opt_state (device, 8 doubles): [0] lr  [1] step  [2] grad sum-sq (scratch, must start at 0)
 *                                  [3] clip coef  [4] step size lr/(1-b1^t)  [5] sqrt(1-b2^t)  [6] grad norm
 *                                  [7] block ticket of the sum-of-squares launch (scratch, must start at 0)
 *   Two launches: sum of squares (its last block derives the scalars), then clip + Adam + shadow refresh + zeroing.
 *   loss_log[log_slot*8 + {0..7}] = {value loss, surrogate loss, lin-vel loss, kl, entropy, lr, grad norm, clip coef}
 *   (means over rows_global rows) -- read back by the host once per update.
 */
#define HG_STATS_SLOTS 8
#define HG_MAX_SHADOWS 16
typedef struct HgBf16Shadow {   /* refresh of one bf16 weight copy after the step */
  int64_t param_offset;         /* offset of the [rows, cols] fp32 weight in the arena */
  int32_t rows, cols;
  void* dst;                    /* bf16 [rows, ld] */
  int64_t ld;
  /* 0 = plain copy.  Otherwise the cols are frames of src_frame columns and land at a pitch of dst_frame columns:
   * column f*src_frame + c -> dst column f*dst_frame + c (first-layer weights for the frame-pitched history ring). */
  int32_t src_frame, dst_frame;
} HgBf16Shadow;

typedef struct HgAdamArgs {
  float* params; float* grads; float* exp_avg; float* exp_avg_sq;
  int64_t count;
  double* opt_state;
  double* loss_log; int32_t log_slot;
  int64_t rows_global;          /* rows in the minibatch over all ranks (kl / loss means) */
  float beta1, beta2, eps, max_grad_norm;
  int32_t adaptive; float desired_kl; float lr_min, lr_max;
  const HgBf16Shadow* shadows_host; int32_t num_shadows;
  int32_t zero_grads;           /* clear the gradient arena after the step */
} HgAdamArgs;

int hg_clip_adam_step(const HgAdamArgs* args_host, void* stream);

/* fp32 [rows, cols] (row stride cols) -> bf16 [rows, ld], zero pad; used once at start-up.  src_frame / dst_frame as in
 * HgBf16Shadow (0, 0 = plain). */
int hg_cast_weight_bf16(const float* src, int rows, int cols, void* dst, int64_t ld, int src_frame, int dst_frame, void* stream);

/* ------------------------------------------------------------------------------------------
 * Env tensor math: one fused launch per env step.  Replaces, in humanoid/envs/custom/humanoid_env.py,
 * post_physics_step's tensor work (:780-806): base-frame quantities (:784-788), check_termination
 * (:811-816), compute_reward with the 22 active _reward_* terms (:889-907,:1170-1437), the buffer and
 * history zeroing of reset_idx (:1132-1139,:1160-1163), compute_observations incl. compute_ref_state,
 * noise and the 15/3-frame history stacks (:717-744,:818-887), the state hand-over (:802-806) and the
 * +-18 clips of step (:654-657).
 */
#define HG_NUM_REWARDS 22
/* alphabetical order of the active reward terms (utils/helpers.py:43-58 walks dir()) */
enum HgReward {
  HG_R_ACTION_SMOOTHNESS = 0, HG_R_BASE_ACC, HG_R_BASE_HEIGHT, HG_R_COLLISION, HG_R_DEFAULT_JOINT_POS,
  HG_R_DOF_ACC, HG_R_DOF_VEL, HG_R_FEET_AIR_TIME, HG_R_FEET_CLEARANCE, HG_R_FEET_CONTACT_FORCES,
  HG_R_FEET_CONTACT_NUMBER, HG_R_FEET_DISTANCE, HG_R_FOOT_SLIP, HG_R_JOINT_POS, HG_R_KNEE_DISTANCE,
  HG_R_LOW_SPEED, HG_R_ORIENTATION, HG_R_TORQUES, HG_R_TRACK_VEL_HARD, HG_R_TRACKING_ANG_VEL,
  HG_R_TRACKING_LIN_VEL, HG_R_VEL_MISMATCH_EXP
};

typedef struct HgEnvConfig {
  int32_t num_envs, D, leg_offset, num_bodies;
  int32_t feet[2], knees[2];
  int32_t num_term_bodies, term_bodies[4];
  int32_t num_pen_bodies, pen_bodies[4];
  int32_t frame_stack, c_frame_stack;
  float dt, cycle_time, max_episode_length;
  float target_joint_pos_scale, base_height_target, min_dist, max_dist, target_feet_height;
  float tracking_sigma, max_contact_force;
  int32_t only_positive_rewards;
  float lin_vel_scale, ang_vel_scale, dof_pos_scale, dof_vel_scale, quat_scale;
  float clip_obs, noise_level;
  float reward_scale_dt[HG_NUM_REWARDS]; /* scale*dt per term; 0 disables the term */
  int32_t phase_mul_reciprocal;  /* 0: phase = len*dt/cycle (torch-CPU semantics); 1: len*dt*(1/cycle) (torch-CUDA) */
  /* HG_ENV_PHASE_REWARD: episode_length += 1, termination, rewards (what precedes reset_idx in post_physics_step);
   * HG_ENV_PHASE_OBS: buffer/history zeroing of reset envs, observations, history, hand-over (what follows it).
   * A real Isaac Gym env launches the two separately around its gym reset calls; synthetic loops pass both. */
  int32_t phases;
  /* Optional compact rigid-body table: rigid_state is [N, num_rigid_bodies, 13] and rigid_feet / rigid_knees index into it
   * (a simulator bridge that exports only the four consumed body rows: 208 B instead of 676 B per env for XBot-L).
   * num_rigid_bodies = 0: rigid_state is [N, num_bodies, 13], indexed by feet / knees.  contact_forces can be compacted
   * without extra fields: num_bodies, feet, term_bodies, pen_bodies then refer to the compact contact table. */
  int32_t num_rigid_bodies, rigid_feet[2], rigid_knees[2];
} HgEnvConfig;
#define HG_ENV_PHASE_REWARD 1
#define HG_ENV_PHASE_OBS 2
#define HG_ENV_PHASE_ALL 3

/* O(1) history ring (reference: the obs_history / critic_history deques, humanoid_env.py:880-887, and
 * their zeroing in reset_idx :1159-1163).  Env-major, time-contiguous: the stacked observation of storage row r of env n
 * is the WINDOW [n, r .. r+frames-1, :] -- a strided view, never re-stacked.  Per env step only the newest frame is
 * written (slot r+frames-1 for row r): three copies,
 *   live32  the history exactly as the reference's deque holds it: a reset zeroes the env's previous frames-1 slots in
 *           place, so live32 windows ARE the reference's obs_buf / privileged_obs_buf (API-visible views);
 *   live16  the same in bf16 at a 16-byte-aligned frame pitch (pad columns zero): the TMA source of the first layer during
 *           the rollout (tensor map base = live16 + r*pitch16, row stride = slots*pitch16, extent = frames*pitch16);
 *   arch32  never zeroed: together with first_valid[r, n] (frames older than that time read as zero) it reproduces
 *           every row of the iteration (hg_ring_gather), after later resets too: the source of the API-visible fp32 rows;
 *   arch16  the same archive in bf16 at the frame pitch: the source of the update's minibatch gather (half the bytes). */
typedef struct HgHistoryRing {
  float* live32;    /* [N, slots, width] */
  void* live16;     /* [N, slots, pitch16] bf16 */
  float* arch32;    /* [N, slots, width] */
  void* arch16;     /* [N, slots, pitch16] bf16 */
  int32_t width;    /* So / Sp */
  int32_t pitch16;  /* width rounded up to 8 */
  int32_t frames;   /* F / Fc */
  int32_t slots;    /* rows + frames - 1 */
} HgHistoryRing;

typedef struct HgEnvBuffers {
  /* Isaac-Gym-shaped inputs */
  const float* root_states;     /* [N,13] */
  const float* dof_state;       /* [N*D,2] interleaved pos,vel */
  const float* contact_forces;  /* [N,B,3] */
  const float* rigid_state;     /* [N,B,13] */
  const float* commands;        /* [N,4] */
  const float* torques;         /* [N,D] */
  const float* rand_push_force; /* [N,3] */
  const float* rand_push_torque;/* [N,3] */
  const float* env_frictions;   /* [N] */
  const float* body_mass;       /* [N] */
  const float* default_dof_pos; /* [D] */
  const float* noise_scale_vec; /* [So] */
  const float* noise;           /* [N,So] N(0,1) draws, or NULL -> Philox(seed, counter) */
  uint64_t seed, counter;
  /* optional: the policy's raw actions [N,D]; when set, the kernel applies clip(+-clip_actions) itself (the clip
   * of step(), humanoid_env.py:633-635) and stores the result in `actions` */
  const float* actions_in;
  float clip_actions;
  /* persistent state, updated in place */
  float* actions;               /* [N,D] in: this step's applied actions; zeroed on reset */
  float* last_actions; float* last_last_actions;   /* [N,D] */
  float* last_dof_vel;          /* [N,D] */
  float* last_root_vel;         /* [N,6] */
  float* feet_air_time;         /* [N,2] */
  uint8_t* last_contacts;       /* [N,2] */
  float* feet_height;           /* [N,2] */
  float* last_feet_z;           /* [N,2] */
  float* ref_dof_pos;           /* [N,D] */
  int64_t* episode_length_buf;  /* [N] */
  float* episode_sums;          /* [HG_NUM_REWARDS, N] */
  /* history in / out: stacked frames oldest->newest; prev and next may NOT alias */
  const float* obs_prev; int64_t obs_prev_ld;   /* [N, F*So] */
  const float* priv_prev; int64_t priv_prev_ld; /* [N, Fc*Sp] */
  float* obs_next; int64_t obs_next_ld;
  float* priv_next; int64_t priv_next_ld;
  /* optional bf16 shadows of obs_next / priv_next (same values, row stride in elements): lets the env write the
   * learner's next storage slot AND its tensor-core shadow directly, so PPO.act needs no ingest launch */
  void* obs16_next; int64_t obs16_next_ld;
  void* priv16_next; int64_t priv16_next_ld;
  /* outputs */
  float* rew_buf;               /* [N] */
  uint8_t* reset_buf;           /* [N] */
  uint8_t* time_out_buf;        /* [N] */
  float* reward_terms;          /* [HG_NUM_REWARDS, N] or NULL (debug/parity) */
  /* optional (all three or none): the learner's transition record for the step these rewards belong to, i.e. what
   * PPO.process_env_step + RolloutStorage.add_transitions would store (ppo.py:129-133, rollout_storage.py:93-94):
   *   store_rewards[n] = rew + store_gamma * store_values[n] * time_out,   store_dones[n] = reset */
  const float* store_values;    /* [N] value estimates of the step that produced the actions */
  float* store_rewards;         /* [N] */
  uint8_t* store_dones;         /* [N] */
  float store_gamma;
  /* optional [N,6]: root_states[:,7:13] as they were BEFORE this step's push (written by hg_command_step).  The
   * reference rotates base_lin_vel / base_ang_vel from the root state before `_post_physics_step_callback` pushes
   * the robots (humanoid_env.py:785-791), while base_acc and last_root_vel read the pushed values (:803,1385-1391).
   * NULL: the base-frame velocities come from root_states (no push this step). */
  const float* root_vel_prepush;
  /* optional history ring (ring_obs.live32 != NULL selects it; obs_prev / obs_next / *16_next are then unused): the
   * newest frames go to slot ring_row + frames - 1, first_valid[ring_row, n] and hist_start[n] are updated */
  HgHistoryRing ring_obs, ring_priv;
  int32_t ring_row;             /* storage row this step produces (1 .. rows-1) */
  int32_t* ring_first_valid;    /* [rows, N]: first valid frame time of each row (a frame's time = the row it is newest in) */
  int32_t* ring_hist_start;     /* [N]: first valid frame time of the env's current episode */
} HgEnvBuffers;

int hg_env_step(const HgEnvConfig* cfg_host, const HgEnvBuffers* buf_host, void* stream);

/* Rows of the iteration out of the history ring (replaces the index gathers of mini_batch_generator,
 * rollout_storage.py:176-189, and the re-stacking of :880-887 for the API-visible [T, N, F*width] view):
 *   out row i = stacked frames of flat row idx[i] = r*N + n (idx NULL: i itself), frames older than first_valid[r, n] zero.
 *   dst_bf16 = 1: bf16 [rows_out, ld_dst] at frame pitch ring->pitch16 (pad columns zero) -- the minibatch operand;
 *   dst_bf16 = 0: fp32 [rows_out, ld_dst] packed (frame pitch = width) -- the reference's layout.
 *   lv_out (optional, fp32 [rows_out, 4]): columns lv_col .. lv_col+2 of the packed row (the lin-vel regression target
 *   critic_obs[:, 53:56], rollout_storage.py:164); dst may then be NULL (targets only). */
int hg_ring_gather(const HgHistoryRing* ring_host, int num_envs, const int32_t* first_valid, const int64_t* idx,
                   int64_t rows_out, void* dst, int64_t ld_dst, int dst_bf16, float* lv_out, int lv_col, void* stream);
/* Start of the next iteration: the window of row `from_row` becomes row 0 (slots [from_row, from_row+frames-1] ->
 * [0, frames-1] in live32 / live16 / arch32 of every ring), hist_start -= from_row, first_valid[0, n] = hist_start[n]. */
int hg_ring_rollover(const HgHistoryRing* rings_host, int num_rings, int num_envs, int from_row, int32_t* hist_start,
                     int32_t* first_valid_row0, void* stream);

/* Pre-physics tensor math of XBotLFreeEnv.step (humanoid_env.py:624-635) and the PD law of _compute_torques
 * (:910-925), the step immediately in front of the hot path, one launch:
 *   a   = (1 - delay[n]) * raw[n,d] + delay[n] * actions[n,d]          (action delay mix, :629)
 *   a  += dyn_rand * noise[n,d] * a                                     (multiplicative noise, :633)
 *   actions[n,d] = clip(a, +-clip_actions)                              (:636)
 *   torques[n,d] = clip(p_gains*(actions*action_scale + default_dof_pos[d] - dof_pos) - d_gains*dof_vel, +-torque_limits[d])
 * delay [N] ~ U[0,1) and noise [N,D] ~ N(0,1) are given, or NULL -> in-kernel Philox4x32-10 (seed, counter).
 * raw_actions NULL skips the action part (a real env re-evaluates only the torque law on each of its 10 physics
 * sub-steps, :639-649).  torques NULL skips the torque part. */
typedef struct HgActionArgs {
  int32_t num_envs, D;
  const float* raw_actions;    /* [N,D] policy output or NULL */
  const float* delay;          /* [N] or NULL */
  const float* noise;          /* [N,D] or NULL */
  uint64_t seed, counter;
  float dyn_rand, clip_actions, action_scale;
  float* actions;              /* [N,D] in: previous applied actions, out: this step's */
  const float* dof_state;      /* [N*D,2] interleaved pos, vel */
  const float* p_gains;        /* [N,D] */
  const float* d_gains;        /* [N,D] */
  const float* default_dof_pos;/* [D] */
  const float* torque_limits;  /* [D] */
  float* torques;              /* [N,D] or NULL */
} HgActionArgs;
int hg_action_torque(const HgActionArgs* args_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mirror-symmetry loss of PPO.update (ppo.py:93-105,197-202; off in the shipped config).
 * A signed permutation matrix P (P[|p_i|][i] = sign(p_i), ppo.py:96-105) is passed as src[i] = |p_i| and
 * sign[i] in {-1, 0, +1} (np.sign(0) = 0 zeroes the column, as in the reference).
 *   hg_permute_cols_bf16: dst[r, i] = sign[i] * src[r, src_col[i]]  = (obs_batch @ obs_perm_mat) on the bf16 rows (:199)
 *   hg_sym_loss:          diff = mu - mirror_act @ act_perm_mat; loss_sum += sum diff^2 (:201-202);
 *                         dmu = scale * diff, dmirror = -scale * diff @ act_perm_mat^T   (scale = 2 sym_coef / (rows A world))
 *   hg_head_backward:     backward of an fp32 output layer out = h W^T + b for a given d/d out: dZ (+)= (dOut W) * elu'(h)
 *                         (bf16, the top hidden layer's pre-activation gradient), gW += dOut^T h, gb += colsum(dOut),
 *                         gb_hidden += colsum(dZ contribution).  n_out <= 16.
 * The mirrored actor forward and the hidden-layer backward reuse the grouped GEMMs and hg_head_forward. */
int hg_permute_cols_bf16(const void* src, int64_t ld_src, int rows, int cols, const int32_t* src_col, const float* sign,
                         void* dst, int64_t ld_dst, void* stream);
int hg_sym_loss(const float* mu, const float* mirror_act, const int32_t* act_src, const float* act_sign, int rows,
                int num_actions, float scale, float* dmu, float* dmirror, double* loss_sum, void* stream);
int hg_head_backward(const void* h, int64_t ld_h, int H, const float* W, int n_out, const float* d_out, int rows,
                     int accumulate, void* dZ, int64_t ld_dz, float* gW, float* gb, float* gb_hidden, void* stream);

/* The tensor math around the reward / observation kernels of post_physics_step, one launch each:
 *
 * HG_CMD_CALLBACK = `_post_physics_step_callback` (humanoid_env.py:1001-1016):
 *   envs with (episode_length + ep_increment) % resampling_steps == 0 get new commands (`_resample_commands`,
 *   :1018-1032: x, y and heading -- or yaw rate when heading_command is 0 -- drawn as scale * u + lo, then x and y are
 *   zeroed where their norm is <= 0.2); with heading_command the yaw-rate command of EVERY env becomes
 *   clip(0.5 * wrap_to_pi(commands[:,3] - heading), -1, 1), heading = atan2 of quat_apply(base_quat, (1,0,0)) (:1007-1010,
 *   utils/math.py:46-49); with push != 0 (`_push_robots`, :665-681, decided by the host from
 *   common_step_counter % push_interval) root_states[:,7:9] = rand_push_force[:,:2] = U(-max_push_vel_xy, +) and
 *   root_states[:,10:13] = rand_push_torque = U(-max_push_ang_vel, +), the pre-push velocities going to root_vel_prepush.
 * HG_CMD_RESET = the `_resample_commands(env_ids)` call of `reset_idx` (:1127) for the envs whose reset_buf is set.
 *
 * u_cmd [N,3] / u_push [N,5]: U[0,1) draws (torch.rand), or NULL -> in-kernel Philox4x32-10 (seed, counter). */
#define HG_CMD_CALLBACK 0
#define HG_CMD_RESET 1
typedef struct HgCommandArgs {
  int32_t num_envs, mode, resampling_steps, ep_increment, heading_command, push;
  float cmd_scale[4], cmd_lo[4];       /* x, y, yaw rate, heading: (float)(hi - lo) and (float)lo of command_ranges */
  float push_vel_scale, push_vel_lo;   /* (float)(2 max_push_vel_xy), (float)(-max_push_vel_xy) */
  float push_ang_scale, push_ang_lo;
  const int64_t* episode_length_buf;   /* [N] (HG_CMD_CALLBACK) */
  const uint8_t* reset_buf;            /* [N] (HG_CMD_RESET) */
  const float* u_cmd;                  /* [N,3] or NULL */
  const float* u_push;                 /* [N,5] or NULL */
  uint64_t seed, counter;
  float* commands;                     /* [N,4] in / out */
  float* root_states;                  /* [N,13] quaternion read; velocities overwritten by a push */
  float* root_vel_prepush;             /* [N,6] out, or NULL */
  float* rand_push_force;              /* [N,3] */
  float* rand_push_torque;             /* [N,3] */
} HgCommandArgs;
int hg_command_step(const HgCommandArgs* args_host, void* stream);

/* extras["episode"]["rew_<term>"] of reset_idx (humanoid_env.py:1141-1143): for each of the K rows of episode_sums [K,N],
 * mean over the envs with reset_buf set, divided by max_episode_length_s.  Launch it between the HG_ENV_PHASE_REWARD and
 * HG_ENV_PHASE_OBS launches (the second one zeroes the sums).  count_out[0] = number of reset envs; when it is 0 the
 * means are left untouched (reset_idx returns early, :1113-1114).  Fixed-order double-precision reduction. */
int hg_episode_means(const float* episode_sums, const uint8_t* reset_buf, int num_terms, int num_envs,
                     float max_episode_length_s, float* means_out, int32_t* count_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HGB200_H_ */
