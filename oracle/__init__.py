"""CPU oracle for the humanoid-gym PPO hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`humanoid_gym_with_comments_b200/`) may import this package: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and there only as the checker / reported CPU baseline.

Contents
  isaac_math.py   restatement of the 4 Isaac Gym `torch_utils` helpers on the path
                  (third-party, NOT under /root/reference -> "parity unpinned",
                  cross-checked against reference `scripts/sim2sim.py:57-89`).
  ppo_oracle.py   restatement of humanoid/algo/ppo/{ppo,actor_critic,rollout_storage}.py
  env_oracle.py   restatement of humanoid/envs/custom/humanoid_env.py hot-path math
  ref_harness.py  loads the REAL reference from /root/reference (this container only)
  make_golden.py  runs reference + oracle on seeded inputs, asserts they agree,
                  writes tests/golden/*.npz

The oracle is pinned: `make_golden.py` executed the unmodified reference here and
the committed fixtures hold the reference's outputs (see tests/golden/MANIFEST.json).
"""
