"""Device-side timeline of the rollout kernels (env step, layer-1 GEMM, tail) from a trace build:
    ALL=1 tools/build_variant.sh trace -DHG_TRACE
    HG_LIB_PATH=$PWD/variants/libhgb200_trace.so python tools/trace_rollout.py
Every traced kernel writes {globaltimer, tag, block} records (csrc/hg_common.cuh: trace()); this prints, per launch of a
rollout step, when each phase was reached (ns after the step's first record): min / median / max over the blocks."""
import contextlib, ctypes as C, io, os, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from humanoid_gym_with_comments_b200 import _lib
from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig

N, T = int(os.environ.get("HG_ENVS", 4096)), int(os.environ.get("HG_T", 60))
STEPS = int(os.environ.get("HG_TRACE_STEPS", 4))
dev = "cuda:0"
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])
alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
          entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, schedule="adaptive", desired_kl=0.01, device=dev)
alg.init_storage(N, T, [705], [219], [12])
env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), dev, torch_rng_noise=False)
env.bind_storage(alg.storage, gamma=alg.gamma)
alg.torch_rng = False
obs, cobs = env.get_observations(), env.get_privileged_observations()


def rollout(steps):
    global obs, cobs
    with torch.inference_mode():
        for i in range(steps):
            a = alg.act(obs, cobs)
            obs, cobs, rew, dones, infos = env.step(a)
            alg.process_env_step(rew, dones, infos)


for it in range(2):                      # warm-up iterations (plans, caches)
    rollout(T)
    with torch.inference_mode():
        alg.compute_returns(cobs)
    alg.update()
rollout(8)
torch.cuda.synchronize()
lib = _lib.load()
TB = 512
table = torch.zeros(4 * 64 * 4 * TB, dtype=torch.int64, device=dev)
lib.hg_trace_set.restype = C.c_int
lib.hg_trace_set.argtypes = [C.c_void_p]
NAMES = {0x100: "env entry", 0x101: "env after pdl_wait", 0x105: "env staged (after barrier)", 0x102: "env after role-math barrier",
         0x103: "env rewards stored", 0x104: "env hand-over done",
         0x200: "gemm entry", 0x201: "gemm set-up done (after pdl_wait)", 0x202: "gemm first TMA issued",
         0x203: "gemm first operands landed", 0x204: "gemm tile MMAs committed", 0x205: "gemm epilogue sees accumulator",
         0x206: "gemm epilogue tile done", 0x207: "gemm end", 0x300: "tail entry", 0x301: "tail set-up done (after pdl_wait)",
         0x302: "tail first operands landed", 0x303: "tail layer-2 MMAs committed", 0x304: "tail epilogue sees acc1",
         0x305: "tail h2 parked", 0x306: "tail W3 + h2 ready", 0x307: "tail epilogue sees acc2", 0x309: "tail dots stored",
         0x30a: "tail head finished", 0x308: "tail end"}
for wv in range(4):
    NAMES[0x110 + wv] = f"env role math done warp {wv}"
    NAMES[0x120 + wv] = f"env frames written warp {wv}"
SUBNAME = {0x204: "tile", 0x205: "tile", 0x206: "tile", 0x307: "net", 0x308: "net", 0x309: "net", 0x30a: "net"}
for s_ in range(STEPS):
    table.zero_()
    torch.cuda.synchronize()
    assert lib.hg_trace_set(table.data_ptr()) == 0
    rollout(1)
    torch.cuda.synchronize()
    assert lib.hg_trace_set(None) == 0
    rollout(2)                                     # untraced steps in between (the traced step starts behind a normal one)
    torch.cuda.synchronize()
    h = table.cpu().view(4, 64, 4, TB)
    t0 = int(h[h > 0].min())
    print(f"--- traced step {s_}: {int(h.max()) - t0} ns from the first to the last record")
    rows = []
    for tag in NAMES:
        k, pt = (tag >> 8) & 3, tag & 63
        for sub in range(4):
            v = h[k, pt, sub]
            v = sorted((v[v > 0] - t0).tolist())
            if v:
                label = NAMES[tag] + (f" ({SUBNAME[tag]} {sub})" if tag in SUBNAME else "")
                rows.append((v[len(v) // 2], f"  {label:46s} n={len(v):4d}  min {v[0]:7d}  med {v[len(v) // 2]:7d}  max {v[-1]:7d}"))
    for _, line in sorted(rows):
        print(line)
