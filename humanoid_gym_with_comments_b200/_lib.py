"""ctypes binding of libhgb200.so (C ABI declared in include/hgb200.h) + the in-tree nvcc build.

The library is the product: if it is missing or a call fails, we raise -- there is no CPU or torch
fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("HG_LIB_PATH") or os.path.join(HERE, "libhgb200.so")   # HG_LIB_PATH: kernel-variant experiments (tools/build_variant.sh)
BUILD_DIR = os.path.join(HERE, "csrc", "_build")
SOURCES = ["hg_runtime.cu", "hg_gemm.cu", "hg_gae.cu", "hg_misc.cu", "hg_heads.cu", "hg_optim.cu", "hg_env.cu", "hg_env_aux.cu", "hg_sym.cu", "hg_ring.cu"]
EXTRA_FLAGS = {"hg_env.cu": ["-fmad=false"], "hg_env_aux.cu": ["-fmad=false"]}
if os.environ.get("HG_EPI_WARPS"):
    EXTRA_FLAGS["hg_gemm.cu"] = ["-DHG_EPI_WARPS=" + os.environ["HG_EPI_WARPS"]]
if os.environ.get("HG_GEMM_DEFS"):
    EXTRA_FLAGS.setdefault("hg_gemm.cu", []).extend(os.environ["HG_GEMM_DEFS"].split())
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]

HG_MAX_GROUP = 4
HG_NUM_REWARDS = 22
HG_STATS_SLOTS = 8
HG_MAX_SHADOWS = 16
ENV_PHASE_REWARD, ENV_PHASE_OBS, ENV_PHASE_ALL = 1, 2, 3
EPI_STORE_F32, EPI_BIAS_ELU_BF16, EPI_DELU_BF16, EPI_ATOMIC_F32, EPI_BIAS_BF16, EPI_BIAS_ELU_HEAD = range(6)
HEAD_NONE, HEAD_POLICY, HEAD_LINEAR = range(3)

REWARD_NAMES = ["action_smoothness", "base_acc", "base_height", "collision", "default_joint_pos", "dof_acc", "dof_vel",
                "feet_air_time", "feet_clearance", "feet_contact_forces", "feet_contact_number", "feet_distance",
                "foot_slip", "joint_pos", "knee_distance", "low_speed", "orientation", "torques", "track_vel_hard",
                "tracking_ang_vel", "tracking_lin_vel", "vel_mismatch_exp"]


class HgError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "hgb200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libhgb200.so in-tree (cudart static, no libcuda link)."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    # One builder at a time (torchrun starts every rank at once): whoever gets the lock builds, the others wait and
    # then find the library fresh.  Objects are rebuilt only when their source (or a header) is newer.
    import fcntl
    with open(os.path.join(BUILD_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            return _build_locked(verbose, force)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool, force: bool) -> str:
    nvcc = _nvcc()
    tag = f".{os.getpid()}"
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "hgb200.h"))
    flags_sig = " ".join(NVCC_FLAGS) + repr(sorted(EXTRA_FLAGS.items()))

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        sig = obj + ".flags"
        deps = [os.path.join(CSRC, src), *headers]
        if (not force and os.path.exists(obj) and os.path.exists(sig) and open(sig).read() == flags_sig
                and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps)):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise HgError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr, file=sys.stderr)
        open(sig, "w").write(flags_sig)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + tag + ".tmp"
    r = subprocess.run([nvcc, "-shared", "-cudart", "static", "-o", tmp, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise HgError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


# ------------------------------------------------------------------------------------------ structs
class HgGemmProblem(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("lda", C.c_int64), ("ldb", C.c_int64),
                ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32), ("M", C.c_int32), ("N", C.c_int32),
                ("K", C.c_int32), ("k_splits", C.c_int32), ("out", C.c_void_p), ("ldo", C.c_int64),
                ("bias", C.c_void_p), ("act", C.c_void_p), ("ldact", C.c_int64), ("colsum", C.c_void_p),
                ("out_frame_src", C.c_int32), ("out_frame_dst", C.c_int32)]


class HgFusedHead(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_out", C.c_int32), ("W", C.c_void_p), ("b", C.c_void_p), ("out", C.c_void_p),
                ("std", C.c_void_p), ("actions", C.c_void_p), ("sigma", C.c_void_p), ("log_prob", C.c_void_p)]


class HgHeadNet(C.Structure):
    _fields_ = [("h", C.c_void_p), ("ld_h", C.c_int64), ("H", C.c_int32), ("n_out", C.c_int32),
                ("W", C.c_void_p), ("b", C.c_void_p), ("gW", C.c_void_p), ("gb", C.c_void_p),
                ("dZ", C.c_void_p), ("ld_dz", C.c_int64), ("gb_hidden", C.c_void_p)]


class HgLossArgs(C.Structure):
    _fields_ = [("actor", HgHeadNet), ("critic", HgHeadNet), ("linvel", HgHeadNet),
                ("rows", C.c_int32), ("A", C.c_int32), ("idx", C.c_void_p),
                ("actions", C.c_void_p), ("old_mu", C.c_void_p), ("old_sigma", C.c_void_p),
                ("old_log_prob", C.c_void_p), ("advantages", C.c_void_p), ("returns", C.c_void_p),
                ("old_values", C.c_void_p), ("critic_obs", C.c_void_p), ("ld_cobs", C.c_int64),
                ("lv_col", C.c_int32), ("std", C.c_void_p), ("g_std", C.c_void_p),
                ("clip_param", C.c_float), ("value_loss_coef", C.c_float), ("entropy_coef", C.c_float),
                ("lin_vel_coef", C.c_float), ("use_clipped_value_loss", C.c_int32), ("inv_world", C.c_float),
                ("stats", C.c_void_p)]


class HgBf16Shadow(C.Structure):
    _fields_ = [("param_offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32), ("dst", C.c_void_p),
                ("ld", C.c_int64), ("src_frame", C.c_int32), ("dst_frame", C.c_int32)]


class HgAdamArgs(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("count", C.c_int64), ("opt_state", C.c_void_p), ("loss_log", C.c_void_p), ("log_slot", C.c_int32),
                ("rows_global", C.c_int64), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("max_grad_norm", C.c_float), ("adaptive", C.c_int32), ("desired_kl", C.c_float),
                ("lr_min", C.c_float), ("lr_max", C.c_float), ("shadows_host", C.POINTER(HgBf16Shadow)),
                ("num_shadows", C.c_int32), ("zero_grads", C.c_int32)]


class HgEnvConfig(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("D", C.c_int32), ("leg_offset", C.c_int32), ("num_bodies", C.c_int32),
                ("feet", C.c_int32 * 2), ("knees", C.c_int32 * 2),
                ("num_term_bodies", C.c_int32), ("term_bodies", C.c_int32 * 4),
                ("num_pen_bodies", C.c_int32), ("pen_bodies", C.c_int32 * 4),
                ("frame_stack", C.c_int32), ("c_frame_stack", C.c_int32),
                ("dt", C.c_float), ("cycle_time", C.c_float), ("max_episode_length", C.c_float),
                ("target_joint_pos_scale", C.c_float), ("base_height_target", C.c_float), ("min_dist", C.c_float),
                ("max_dist", C.c_float), ("target_feet_height", C.c_float), ("tracking_sigma", C.c_float),
                ("max_contact_force", C.c_float), ("only_positive_rewards", C.c_int32),
                ("lin_vel_scale", C.c_float), ("ang_vel_scale", C.c_float), ("dof_pos_scale", C.c_float),
                ("dof_vel_scale", C.c_float), ("quat_scale", C.c_float), ("clip_obs", C.c_float),
                ("noise_level", C.c_float), ("reward_scale_dt", C.c_float * HG_NUM_REWARDS),
                ("phase_mul_reciprocal", C.c_int32), ("phases", C.c_int32),
                ("num_rigid_bodies", C.c_int32), ("rigid_feet", C.c_int32 * 2), ("rigid_knees", C.c_int32 * 2)]


class HgHistoryRing(C.Structure):
    _fields_ = [("live32", C.c_void_p), ("live16", C.c_void_p), ("arch32", C.c_void_p), ("arch16", C.c_void_p), ("width", C.c_int32),
                ("pitch16", C.c_int32), ("frames", C.c_int32), ("slots", C.c_int32)]


class HgEnvBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("root_states", "dof_state", "contact_forces", "rigid_state", "commands", "torques",
                 "rand_push_force", "rand_push_torque", "env_frictions", "body_mass", "default_dof_pos",
                 "noise_scale_vec", "noise")] + \
               [("seed", C.c_uint64), ("counter", C.c_uint64), ("actions_in", C.c_void_p), ("clip_actions", C.c_float)] + \
               [(n, C.c_void_p) for n in
                ("actions", "last_actions", "last_last_actions", "last_dof_vel", "last_root_vel", "feet_air_time",
                 "last_contacts", "feet_height", "last_feet_z", "ref_dof_pos", "episode_length_buf", "episode_sums")] + \
               [("obs_prev", C.c_void_p), ("obs_prev_ld", C.c_int64), ("priv_prev", C.c_void_p),
                ("priv_prev_ld", C.c_int64), ("obs_next", C.c_void_p), ("obs_next_ld", C.c_int64),
                ("priv_next", C.c_void_p), ("priv_next_ld", C.c_int64),
                ("obs16_next", C.c_void_p), ("obs16_next_ld", C.c_int64),
                ("priv16_next", C.c_void_p), ("priv16_next_ld", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("rew_buf", "reset_buf", "time_out_buf", "reward_terms", "store_values",
                                          "store_rewards", "store_dones")] + [("store_gamma", C.c_float),
                                                                              ("root_vel_prepush", C.c_void_p)] + \
               [("ring_obs", HgHistoryRing), ("ring_priv", HgHistoryRing), ("ring_row", C.c_int32),
                ("ring_first_valid", C.c_void_p), ("ring_hist_start", C.c_void_p)]


class HgActionArgs(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("D", C.c_int32), ("raw_actions", C.c_void_p), ("delay", C.c_void_p),
                ("noise", C.c_void_p), ("seed", C.c_uint64), ("counter", C.c_uint64), ("dyn_rand", C.c_float),
                ("clip_actions", C.c_float), ("action_scale", C.c_float), ("actions", C.c_void_p),
                ("dof_state", C.c_void_p), ("p_gains", C.c_void_p), ("d_gains", C.c_void_p),
                ("default_dof_pos", C.c_void_p), ("torque_limits", C.c_void_p), ("torques", C.c_void_p)]


class HgCommandArgs(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("num_envs", "mode", "resampling_steps", "ep_increment", "heading_command", "push")] + \
               [("cmd_scale", C.c_float * 4), ("cmd_lo", C.c_float * 4), ("push_vel_scale", C.c_float),
                ("push_vel_lo", C.c_float), ("push_ang_scale", C.c_float), ("push_ang_lo", C.c_float),
                ("episode_length_buf", C.c_void_p), ("reset_buf", C.c_void_p), ("u_cmd", C.c_void_p), ("u_push", C.c_void_p),
                ("seed", C.c_uint64), ("counter", C.c_uint64), ("commands", C.c_void_p), ("root_states", C.c_void_p),
                ("root_vel_prepush", C.c_void_p), ("rand_push_force", C.c_void_p), ("rand_push_torque", C.c_void_p)]


CMD_CALLBACK, CMD_RESET = 0, 1

_P = C.c_void_p
_SIGNATURES = {
    "hg_last_error": (C.c_char_p, []),
    "hg_abi_version": (C.c_int, []),
    "hg_launch_count": (C.c_int64, []),
    "hg_gemm_grouped": (C.c_int, [C.POINTER(HgGemmProblem), C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "hg_gemm_plan_create": (C.c_int, [C.POINTER(HgGemmProblem), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "hg_gemm_plan_run": (C.c_int, [C.c_void_p, _P]),
    "hg_gemm_plan_destroy": (None, [C.c_void_p]),
    "hg_gemm_plan_create_heads": (C.c_int, [C.POINTER(HgGemmProblem), C.POINTER(HgFusedHead), C.c_int, C.POINTER(C.c_void_p)]),
    "hg_gemm_plan_run_heads": (C.c_int, [C.c_void_p, _P, C.c_uint64, C.c_uint64, _P]),
    "hg_rollout_forward": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, _P, C.c_uint64, C.c_uint64, _P]),
    "hg_gemm_plan_create_tail": (C.c_int, [C.POINTER(HgGemmProblem), C.POINTER(HgGemmProblem), C.POINTER(HgFusedHead), C.c_int,
                                           C.POINTER(C.c_void_p)]),
    "hg_gemm_plan_create_step": (C.c_int, [C.POINTER(HgGemmProblem), C.POINTER(HgGemmProblem), C.POINTER(HgGemmProblem),
                                           C.POINTER(HgFusedHead), C.c_int, C.POINTER(C.c_void_p)]),
    "hg_gae_returns": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P]),
    "hg_ingest_rows": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, C.c_int64, _P, C.c_int64, _P]),
    "hg_gather_rows_bf16": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, _P]),
    "hg_policy_head_rollout": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, _P, _P, C.c_int64, C.c_int, _P, _P, _P,
                                         C.c_uint64, C.c_uint64, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "hg_gaussian_sample": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_int, C.c_int, _P, _P, _P]),
    "hg_gaussian_log_prob": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "hg_value_head": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, C.c_int, _P, _P]),
    "hg_head_forward": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P]),
    "hg_store_rewards_dones": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int, _P, _P, _P]),
    "hg_ppo_head_loss": (C.c_int, [C.POINTER(HgLossArgs), _P]),
    "hg_clip_adam_step": (C.c_int, [C.POINTER(HgAdamArgs), _P]),
    "hg_cast_weight_bf16": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "hg_ring_gather": (C.c_int, [C.POINTER(HgHistoryRing), C.c_int, _P, _P, C.c_int64, _P, C.c_int64, C.c_int, _P, C.c_int, _P]),
    "hg_ring_rollover": (C.c_int, [C.POINTER(HgHistoryRing), C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "hg_env_step": (C.c_int, [C.POINTER(HgEnvConfig), C.POINTER(HgEnvBuffers), _P]),
    "hg_action_torque": (C.c_int, [C.POINTER(HgActionArgs), _P]),
    "hg_permute_cols_bf16": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P]),
    "hg_sym_loss": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "hg_head_backward": (C.c_int, [_P, C.c_int64, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int64, _P, _P, _P, _P]),
    "hg_command_step": (C.c_int, [C.POINTER(HgCommandArgs), _P]),
    "hg_episode_means": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
}
EXPORTED_SYMBOLS = sorted(_SIGNATURES)

_lib = None


def load(auto_build: bool = True) -> C.CDLL:
    """dlopen libhgb200.so (building it first if sources are newer) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if auto_build and _stale():
        build()
    if not os.path.exists(LIB_PATH):
        raise HgError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.hg_abi_version() != 8:
        raise HgError("libhgb200.so ABI mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise HgError(f"libhgb200 call failed ({rc}): {load().hg_last_error().decode()}")


def require_cuda(*tensors) -> None:
    """No CPU fallback: every tensor handed to the kernels must live on a CUDA device."""
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise HgError("humanoid_gym_with_comments_b200 runs on CUDA (sm_100a) only; got a tensor on " + str(t.device))


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


# NVTX ranges per phase (the reference has no profiler hooks): HG_NVTX=1 turns them on for nsys / ncu --nvtx runs;
# off by default so that the enqueue path pays nothing.
NVTX = os.environ.get("HG_NVTX", "0") == "1"


class nvtx_range:
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


_raw_stream = None


def current_stream() -> int:
    """cudaStream_t of torch's current stream on the current device (fast path: one C call)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream
