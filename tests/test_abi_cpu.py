"""CPU tests (no GPU): the C-ABI library builds for sm_100a, loads, exports every symbol include/hgb200.h
declares (no compute calls), the ctypes mirrors match the header's struct layout, and the host-side classes keep
the reference's API surface and refuse to run without CUDA."""
import contextlib
import ctypes as C
import inspect
import io
import os
import re
import subprocess

import pytest
import torch

from humanoid_gym_with_comments_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hgb200.h")


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hgb200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes prototypes and header disagree"
    assert lib.hg_abi_version() == 8
    assert lib.hg_launch_count() == 0            # nothing launched on a CPU box


def test_every_entry_point_is_documented_against_the_reference_interface():
    """INTEGRATION.md names, for every exported entry point, the reference interface it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    # `hg_gemm_plan_{create,run,destroy}` style rows spell several names at once
    for m in re.finditer(r"(hg_[a-z0-9_]*)\{([a-z0-9_,]+)\}", doc):
        doc += " " + " ".join(m.group(1) + part for part in m.group(2).split(","))
    missing = [f for f in header_functions() if f not in doc]
    assert not missing, missing


def test_library_contains_sm100a_tensor_core_code():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):     # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic


def test_struct_layouts_match_header_sizes():
    """Compile a 10-line C program against the header and compare sizeof() with the ctypes mirrors."""
    src = '#include "hgb200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(HgGemmProblem),sizeof(HgHeadNet),sizeof(HgLossArgs),sizeof(HgBf16Shadow),sizeof(HgAdamArgs),' \
          'sizeof(HgEnvConfig),sizeof(HgEnvBuffers));printf("%zu %zu %zu\\n",sizeof(HgFusedHead),sizeof(HgCommandArgs),' \
          'sizeof(HgActionArgs));printf("%zu\\n",sizeof(HgHistoryRing));return 0;}\n'
    tmp = os.path.join(_lib.BUILD_DIR, "abi_probe")
    os.makedirs(_lib.BUILD_DIR, exist_ok=True)
    open(tmp + ".c", "w").write(src)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), tmp + ".c", "-o", tmp], check=True)
    sizes = [int(x) for x in subprocess.run([tmp], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [_lib.HgGemmProblem, _lib.HgHeadNet, _lib.HgLossArgs, _lib.HgBf16Shadow, _lib.HgAdamArgs, _lib.HgEnvConfig,
               _lib.HgEnvBuffers, _lib.HgFusedHead, _lib.HgCommandArgs, _lib.HgActionArgs, _lib.HgHistoryRing]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.hg_gemm_grouped(None, 0, 0, 128, 1, None) == -1
    assert b"num_problems" in lib.hg_last_error()
    assert lib.hg_gae_returns(None, None, None, None, 0.9, 0.9, 1, 1, None, None, None, None) == -1
    assert lib.hg_command_step(None, None) == -1
    assert lib.hg_episode_means(None, None, 22, 16, 24.0, None, None, None) == -1


def test_reference_api_surface():
    from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic, RolloutStorage
    sig = inspect.signature(PPO.__init__)
    assert list(sig.parameters)[1:] == ["actor_critic", "num_learning_epochs", "num_mini_batches", "clip_param", "gamma", "lam",
                                        "value_loss_coef", "entropy_coef", "learning_rate", "max_grad_norm",
                                        "use_clipped_value_loss", "schedule", "desired_kl", "device", "sym_loss",
                                        "obs_permutation", "act_permutation", "frame_stack", "sym_coef", "base_lin_vel_coef"]
    assert sig.parameters["gamma"].default == 0.998 and sig.parameters["lam"].default == 0.95
    sig = inspect.signature(ActorCritic.__init__)
    assert list(sig.parameters)[1:9] == ["num_actor_obs", "num_critic_obs", "num_actions", "actor_hidden_dims",
                                         "critic_hidden_dims", "base_lin_vel_hidden_dims", "init_noise_std", "activation"]
    assert list(inspect.signature(RolloutStorage.__init__).parameters)[1:] == [
        "num_envs", "num_transitions_per_env", "obs_shape", "privileged_obs_shape", "actions_shape", "device"]
    for m in ("init_storage", "test_mode", "train_mode", "act", "process_env_step", "compute_returns", "update"):
        assert callable(getattr(PPO, m))
    for m in ("act", "get_actions_log_prob", "act_inference", "evaluate", "update_distribution", "reset", "forward"):
        assert callable(getattr(ActorCritic, m))
    for m in ("add_transitions", "clear", "compute_returns", "get_statistics", "mini_batch_generator"):
        assert callable(getattr(RolloutStorage, m))


def test_actor_critic_state_dict_and_init_match_oracle():
    from humanoid_gym_with_comments_b200.algo.ppo import ActorCritic
    from oracle import ppo_oracle as po
    buf = io.StringIO()
    torch.manual_seed(0)
    with contextlib.redirect_stdout(buf):
        ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128], foo=1)
    assert "unexpected arguments" in buf.getvalue() and "Actor MLP" in buf.getvalue()
    torch.manual_seed(0)
    p = po.init_params(705, 219, 12, (512, 256, 128), (768, 256, 128), (128, 128), 1.0)
    sd = ac.state_dict()
    assert list(sd.keys()) == po.state_dict_order(p)
    assert all(torch.equal(sd[k], p[k]) for k in sd)
    with pytest.raises(NotImplementedError):
        ac.forward()


def test_no_cpu_fallback():
    from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic, RolloutStorage
    from humanoid_gym_with_comments_b200.envs import XBotEnvMath, XBotMathConfig
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(47, 73, 12, [64, 32], [64, 32], [32])
    with pytest.raises(_lib.HgError):
        PPO(ac, device="cpu")
    with pytest.raises(_lib.HgError):
        RolloutStorage(4, 2, [47], [73], [12], device="cpu")
    with pytest.raises(_lib.HgError):
        XBotEnvMath(XBotMathConfig(num_envs=4), "cpu")
    with pytest.raises(_lib.HgError):
        ac.act_inference(torch.zeros(4, 47))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "humanoid_gym_with_comments_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_env_config_reward_order_and_shapes():
    from humanoid_gym_with_comments_b200.envs import XBotMathConfig
    from oracle import env_oracle as eo
    assert _lib.REWARD_NAMES == eo.xbot_spec(12).active_rewards()
    for D in (12, 18):
        cfg, spec = XBotMathConfig.for_dof(D), eo.xbot_spec(D)
        assert cfg.num_observations == spec.num_obs and cfg.num_privileged_obs == spec.num_priv
        assert torch.equal(cfg.noise_scale_vec(), spec.noise_scale_vec())
        assert cfg.feet_indices == spec.feet and cfg.knee_indices == spec.knees
