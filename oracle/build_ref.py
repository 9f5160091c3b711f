"""Recipe for `oracle/_ref/`: installs the UNMODIFIED reference files of the hot path (and their import closure) from
/root/reference into the git-ignored `oracle/_ref/` so that they travel to the GPU box with the snapshot.
TEST / BASELINE INFRASTRUCTURE ONLY -- nothing in the product package reads this directory.

    python -m oracle.build_ref            # copies, verifies, writes oracle/_ref/MANIFEST.json

The reference is pure Python, so "building" it is a byte-for-byte copy; `MANIFEST.json` records the sha256
of every file so a maintainer can check that what ran on the box IS the reference.  What it enables there:
  * `bench.py --impl reference`: the reference's own PPO / ActorCritic / RolloutStorage (humanoid/algo/ppo) and the
    reference's env math text (humanoid/envs/custom/humanoid_env.py under the stub isaacgym of oracle/ref_harness.py)
    on the box's host cores (`cpu_baseline.kind = "reference"`),
  * the same code on cuda:0 (`reference_torch_cuda` in the product line: the denominator of north_star's >= 50x),
  * `tests/test_runner_gpu.py`: the unmodified `OnPolicyRunner` driving the B200 classes.
Nothing is copied into git history: `oracle/_ref/` is listed in .gitignore (not in .gpurunignore).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("HG_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")

# the hot path's files + exactly what they import from the reference tree
FILES = [
    "humanoid/__init__.py",
    "humanoid/algo/__init__.py",
    "humanoid/algo/vec_env.py",
    "humanoid/algo/ppo/__init__.py",
    "humanoid/algo/ppo/ppo.py",
    "humanoid/algo/ppo/actor_critic.py",
    "humanoid/algo/ppo/rollout_storage.py",
    "humanoid/algo/ppo/on_policy_runner.py",
    "humanoid/envs/custom/humanoid_env.py",
    "humanoid/envs/custom/humanoid_config.py",
    "humanoid/envs/base/base_config.py",
    "humanoid/envs/base/base_task.py",
    "humanoid/envs/base/legged_robot_config.py",
    "humanoid/utils/helpers.py",
    "humanoid/utils/math.py",
    "humanoid/utils/terrain.py",
]


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "MANIFEST.json"))


def build(verbose: bool = True) -> bool:
    """Copy the files if the reference tree is present; returns whether oracle/_ref is usable afterwards."""
    if not os.path.isdir(os.path.join(REF_SRC, "humanoid", "algo", "ppo")):
        if verbose:
            print(f"oracle/_ref: {REF_SRC} is absent; " + ("keeping the prebuilt copy" if available() else "nothing installed"))
        return available()
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        assert _sha(src) == _sha(dst), rel
        manifest[rel] = _sha(dst)
    json.dump({"source": "Rengar-Yang/humanoid-gym-with-comments (unmodified files)", "files": manifest},
              open(os.path.join(REF_DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: installed {len(manifest)} unmodified reference files")
    return True


def verify() -> bool:
    """True iff every file of oracle/_ref still has the recorded digest."""
    if not available():
        return False
    m = json.load(open(os.path.join(REF_DST, "MANIFEST.json")))["files"]
    return all(os.path.isfile(os.path.join(REF_DST, rel)) and _sha(os.path.join(REF_DST, rel)) == h for rel, h in m.items())


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok and verify() else 1)
