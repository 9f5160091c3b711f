"""Tabulates tools/sweep_envs.sh output: env-steps/s, GEMM TFLOP/s and roofline fraction, GAE / env-step GB/s per size."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith("{")]
print(f"{'envs/GPU':>9} {'ms/iter':>9} {'M env-steps/s':>14} {'GEMM TF/s':>10} {'frac':>6} {'GAE us':>8} {'GAE GB/s':>9} {'frac':>6} {'env us':>7} {'env GB/s':>9} {'frac':>6}")
for r in rows:
    if "failed" in r:
        print(f"{r['envs_per_gpu']:>9} failed ({r['failed']})")
        continue
    ro, aux = r["roofline"], r.get("hbm_kernels", {})
    g, e = aux.get("gae_returns", {}), aux.get("env_step", {})
    print(f"{r['config']['envs_per_gpu']:>9} {r['ms_per_step']:>9.2f} {r['value'] / 1e6:>14.2f} {ro['achieved']:>10.1f} {ro['frac']:>6.3f} "
          f"{g.get('us', float('nan')):>8.1f} {g.get('gbs', float('nan')):>9.1f} {g.get('frac', float('nan')):>6.3f} "
          f"{e.get('us', float('nan')):>7.1f} {e.get('gbs', float('nan')):>9.1f} {e.get('frac', float('nan')):>6.3f}")
