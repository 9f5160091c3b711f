"""Host-side orchestration of the learner kernels (private to this package).

`FlatParams` owns the flat fp32 arenas (parameters | gradients | Adam moments) that the nn.Parameters of
`ActorCritic` are views of, plus the bf16 shadows of the hidden-layer weights that the tensor-core GEMMs read.
`MLPEngine` launches grouped forward / backward passes over the three MLPs through the C ABI.

Reference being replaced: the torch autograd graph of `humanoid/algo/ppo/actor_critic.py:52-89` (three
nn.Sequential of Linear/ELU) as driven by `ppo.py:116-126` (rollout) and `ppo.py:154-214` (update).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import (EPI_ATOMIC_F32, EPI_BIAS_ELU_BF16, EPI_DELU_BF16, HG_STATS_SLOTS, HgBf16Shadow,
                     HgGemmProblem, check, current_stream, ptr)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# Row pitch of every bf16 operand, in elements: a multiple of 64 (128 bytes) so that each 128-byte row of a TMA box
# is exactly one L2 line.  With a merely 16-byte-aligned pitch (712 for 705 columns) the rows straddle lines and the
# first-layer GEMMs lose ~25 % (measured in round 1 with HG_LD_ALIGN=8 in tools/bench_gemm.py).
LD_ALIGN = 64


class NetSpec:
    """One MLP: list of (weight, bias) Parameters, hidden layers first, output layer last."""

    def __init__(self, name: str, seq: nn.Sequential):
        self.name = name
        self.linears: List[nn.Linear] = [m for m in seq if isinstance(m, nn.Linear)]
        others = [m for m in seq if not isinstance(m, nn.Linear)]
        for m in others:
            if not (isinstance(m, nn.ELU) and m.alpha == 1.0):
                raise NotImplementedError(
                    f"{name}: only nn.ELU(alpha=1) between Linear layers has a CUDA kernel here (got {type(m).__name__})")
        self.in_dim = self.linears[0].in_features
        self.hidden = [l.out_features for l in self.linears[:-1]]
        self.out_dim = self.linears[-1].out_features
        for w in self.hidden:
            if w % 8 != 0:
                raise NotImplementedError(f"{name}: hidden widths must be multiples of 8 (got {w})")
        if not self.hidden or self.hidden[-1] > 256:
            raise NotImplementedError(f"{name}: needs >=1 hidden layer and a last hidden width <= 256")
        self.num_hidden = len(self.hidden)

    def widths(self) -> List[int]:
        """input width followed by the hidden widths: widths()[l] is the K of hidden layer l+1."""
        return [self.in_dim] + self.hidden


class FlatParams:
    """Flat fp32 arenas + bf16 weight shadows for a list of nn.Parameters (in `module.parameters()` order)."""

    def __init__(self, module: nn.Module, nets: Sequence[NetSpec]):
        self.module = module
        self.nets = list(nets)
        self.params: List[nn.Parameter] = list(module.parameters())
        dev = self.params[0].device
        _lib.require_cuda(*self.params)
        self.device = dev
        self.offsets: Dict[int, int] = {}
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise _lib.HgError("parameters must be float32")
            self.offsets[id(p)] = off
            off += round_up(p.numel(), 4)            # 16-byte aligned starts
        self.count = off
        n = self.count + HG_STATS_SLOTS
        self.p = torch.zeros(n, device=dev)
        self.g = torch.zeros(n, device=dev)
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.stats = self.g[self.count:]              # loss statistics ride behind the gradients (one all-reduce)
        with torch.no_grad():
            for p in self.params:
                o = self.offsets[id(p)]
                view = self.p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.g[o:o + p.numel()].view(p.shape)
        # bf16 shadows of the hidden-layer weights, rows padded to 16 bytes
        self.w16: Dict[int, torch.Tensor] = {}
        self._shadow_list: List[HgBf16Shadow] = []
        for net in self.nets:
            for lin in net.linears[:-1]:
                w = lin.weight
                ld = round_up(w.shape[1], LD_ALIGN)
                sh = torch.zeros(w.shape[0], ld, dtype=torch.bfloat16, device=dev)
                self.w16[id(w)] = sh
                self._shadow_list.append(HgBf16Shadow(self.offsets[id(w)], w.shape[0], w.shape[1], sh.data_ptr(), ld, 0, 0))
        # frame-pitched shadows of first-layer weights (history-ring operands), created on demand: (id(w), width, pitch) -> bf16
        self.w16_framed: Dict[tuple, torch.Tensor] = {}
        self._framed_specs: List[tuple] = []
        self._rebuild_shadow_array()
        self.refresh_shadows()

    def _rebuild_shadow_array(self) -> None:
        if len(self._shadow_list) > _lib.HG_MAX_SHADOWS:
            raise NotImplementedError("too many bf16 weight shadows for one optimiser launch")
        self.shadow_array = (HgBf16Shadow * len(self._shadow_list))(*self._shadow_list)

    def framed_shadow(self, w: nn.Parameter, width: int, pitch: int) -> torch.Tensor:
        """bf16 copy of a first-layer weight [out, frames*width] with every frame of `width` columns moved to a pitch of
        `pitch` columns (pad columns zero): the B operand for input rows read straight out of the history ring, whose
        bf16 frames sit at a 16-byte-aligned pitch.  Kept in step with the fp32 master by the optimiser launch."""
        key = (id(w), width, pitch)
        sh = self.w16_framed.get(key)
        if sh is None:
            if w.shape[1] % width != 0:
                raise _lib.HgError(f"first-layer width {w.shape[1]} is not a whole number of {width}-wide frames")
            frames = w.shape[1] // width
            ld = round_up(frames * pitch, LD_ALIGN)
            sh = torch.zeros(w.shape[0], ld, dtype=torch.bfloat16, device=self.device)
            self.w16_framed[key] = sh
            self._framed_specs.append((w, sh, width, pitch))
            self._shadow_list.append(HgBf16Shadow(self.offsets[id(w)], w.shape[0], w.shape[1], sh.data_ptr(), ld, width, pitch))
            self._rebuild_shadow_array()
            check(_lib.load().hg_cast_weight_bf16(ptr(w.data), w.shape[0], w.shape[1], ptr(sh), ld, width, pitch, current_stream()))
        return sh

    def rebind_grads(self, g: torch.Tensor) -> None:
        """Move the gradient arena into `g` (same length; e.g. a symmetric-memory buffer the peers can read for the
        all-reduce): the `.grad` of every parameter becomes a view of it, the loss statistics stay behind the gradients."""
        if g.numel() != self.g.numel() or g.dtype != torch.float32 or g.device != self.g.device:
            raise _lib.HgError("rebind_grads: need a float32 tensor of the arena's length on the arena's device")
        with torch.no_grad():
            g.copy_(self.g)
        self.g = g
        self.stats = g[self.count:]
        for p in self.params:
            o = self.offsets[id(p)]
            p.grad = g[o:o + p.numel()].view(p.shape)

    # -- views ------------------------------------------------------------------------------
    def grad_view(self, p: nn.Parameter) -> torch.Tensor:
        o = self.offsets[id(p)]
        return self.g[o:o + p.numel()].view(p.shape)

    def moment_views(self, p: nn.Parameter):
        o = self.offsets[id(p)]
        return self.m[o:o + p.numel()].view(p.shape), self.v[o:o + p.numel()].view(p.shape)

    def attached(self) -> bool:
        """False if someone re-materialised the parameters (e.g. `.to()` / `load_state_dict(assign=True)`)."""
        p0 = self.params[0]
        return p0.data.data_ptr() == self.p.data_ptr() + 4 * self.offsets[id(p0)]

    def refresh_shadows(self) -> None:
        """Re-cast every bf16 shadow from the fp32 master weights (after load_state_dict / manual edits)."""
        lib = _lib.load()
        st = current_stream()
        for net in self.nets:
            for lin in net.linears[:-1]:
                w = lin.weight
                sh = self.w16[id(w)]
                check(lib.hg_cast_weight_bf16(ptr(w.data), w.shape[0], w.shape[1], ptr(sh), sh.shape[1], 0, 0, st))
        for w, sh, width, pitch in self._framed_specs:
            check(lib.hg_cast_weight_bf16(ptr(w.data), w.shape[0], w.shape[1], ptr(sh), sh.shape[1], width, pitch, st))


class Workspace:
    """Activation / gradient buffers for `rows` rows of every net (bf16)."""

    def __init__(self, nets: Sequence[NetSpec], rows: int, device, with_grads: bool):
        self.rows = rows
        self.h: Dict[str, List[torch.Tensor]] = {}
        self.dz: Dict[str, List[torch.Tensor]] = {}
        for net in nets:
            self.h[net.name] = [torch.empty(rows, w, dtype=torch.bfloat16, device=device) for w in net.hidden]
            if with_grads:
                self.dz[net.name] = [torch.empty(rows, w, dtype=torch.bfloat16, device=device) for w in net.hidden]


class MLPEngine:
    def __init__(self, flat: FlatParams):
        self.flat = flat
        self.lib = _lib.load()
        self.num_sms = torch.cuda.get_device_properties(flat.device).multi_processor_count
        self._all_plans: List[int] = []
        import os
        self.cta_group_override = int(os.environ.get("HG_CTA_GROUP", "0"))

    # -- helpers ----------------------------------------------------------------------------
    @staticmethod
    def _block_n(ns: Sequence[int], kind: str = "fwd", rows: int = 1 << 30) -> int:
        """UMMA N of the accumulator tile, from tools/bench_gemm.py on B200: 256-wide tiles win whenever some
        problem of the group is at least 256 wide (less L2 -> shared-memory traffic per MMA); rollout-sized launches
        keep 128-wide tiles (more tiles to fill the SMs)."""
        if kind == "dx" and rows > 8192:
            return int(os.environ.get("HG_DX_BN", 256)) if max(ns) >= 256 else 128
        if kind == "fwd" and rows <= 8192:
            return int(os.environ.get("HG_ROLLOUT_BN", 128))       # experiment switch (tools/rollout_variants.sh)
        if kind == "dx" or rows <= 8192:
            return 128       # rollout-sized launches: more, smaller tiles fill the 148 SMs better
        return 256 if max(ns) >= 256 else 128

    def _cta_group(self, problems: List[HgGemmProblem], epi: int = EPI_BIAS_ELU_BF16) -> int:
        """CTA pairs (tcgen05.mma.cta_group::2, 256-row tiles) for the forward and weight-gradient launches once
        there are enough rows to fill the machine; the dX launches (K <= 256, epilogue-bound) keep single CTAs
        (tools/bench_gemm.py)."""
        if self.cta_group_override:
            return self.cta_group_override
        if epi == EPI_DELU_BF16:
            return int(os.environ.get("HG_DX_CG", 1))
        if epi == EPI_ATOMIC_F32 and max(p.M for p in problems) <= 128 and os.environ.get("HG_DW_SMALL_CG", "2") == "1":
            return 1         # experiment: weight gradients of 128-row layers leave half of a 256-row pair tile empty
        if max(p.M for p in problems) >= 8192 or max(p.K for p in problems) >= 8192:
            return 2
        return int(os.environ.get("HG_ROLLOUT_CG", 1)) if epi == EPI_BIAS_ELU_BF16 else 1

    def _launch(self, problems: List[HgGemmProblem], epi: int, block_n: int) -> None:
        arr = (HgGemmProblem * len(problems))(*problems)
        check(self.lib.hg_gemm_grouped(arr, len(problems), epi, block_n, self._cta_group(problems, epi), current_stream()))

    def _plan(self, problems: List[HgGemmProblem], epi: int, block_n: int) -> int:
        arr = (HgGemmProblem * len(problems))(*problems)
        handle = C.c_void_p()
        check(self.lib.hg_gemm_plan_create(arr, len(problems), epi, block_n, self._cta_group(problems, epi), C.byref(handle)))
        self._all_plans.append(handle.value)
        return handle.value

    @staticmethod
    def can_fuse_heads(nets: Sequence[NetSpec], n_outs: Sequence[int]) -> bool:
        """The fused last-layer + head launch needs equal depth, last hidden width <= 128 (multiple of 32) and at most
        16 outputs per head."""
        depth = {n.num_hidden for n in nets}
        return (len(depth) == 1 and all(n.hidden[-1] <= 128 and n.hidden[-1] % 32 == 0 for n in nets)
                and all(1 <= o <= 16 for o in n_outs))

    @staticmethod
    def can_fuse_tail(nets: Sequence[NetSpec]) -> bool:
        """Layer 2 + layer 3 + head in one launch (`hg_gemm_plan_create_tail`): one or two nets with three hidden layers,
        second width a multiple of 64 up to 256, third a multiple of 32 up to 128.  HG_ROLLOUT_TAIL=0 keeps two launches."""
        if os.environ.get("HG_ROLLOUT_TAIL", "1") == "0" or not 1 <= len(nets) <= 2:
            return False
        return all(n.num_hidden == 3 and 64 <= n.hidden[1] <= 256 and n.hidden[1] % 64 == 0
                   and 32 <= n.hidden[2] <= 128 and n.hidden[2] % 32 == 0 for n in nets)

    def _tail_plan(self, nets: Sequence[NetSpec], ws: "Workspace", rows: int, heads, layer1=None) -> int:
        """layer1 = the layer-1 problems of the same nets: ONE launch for the whole forward (`hg_gemm_plan_create_step`);
        raises HgError when the shapes do not qualify (the caller then keeps layer 1 and the tail as two launches)."""
        l2, l3, hs = [], [], []
        for net in nets:
            h1 = ws.h[net.name][0]
            lin2, lin3 = net.linears[1], net.linears[2]
            w2, w3 = self.flat.w16[id(lin2.weight)], self.flat.w16[id(lin3.weight)]
            l2.append(HgGemmProblem(A=ptr(h1), B=ptr(w2), lda=h1.stride(0), ldb=w2.stride(0), a_mn_major=0, b_mn_major=0,
                                    M=rows, N=lin2.out_features, K=net.hidden[0], k_splits=1, out=0, ldo=0,
                                    bias=ptr(lin2.bias.data), act=0, ldact=0, colsum=0, out_frame_src=0, out_frame_dst=0))
            l3.append(HgGemmProblem(A=0, B=ptr(w3), lda=0, ldb=w3.stride(0), a_mn_major=0, b_mn_major=0,
                                    M=rows, N=lin3.out_features, K=net.hidden[1], k_splits=1, out=0, ldo=0,
                                    bias=ptr(lin3.bias.data), act=0, ldact=0, colsum=0, out_frame_src=0, out_frame_dst=0))
            hs.append(heads[net.name])
        a2, a3 = (HgGemmProblem * len(l2))(*l2), (HgGemmProblem * len(l3))(*l3)
        harr = (_lib.HgFusedHead * len(hs))(*hs)
        handle = C.c_void_p()
        if layer1 is not None:
            a1 = (HgGemmProblem * len(layer1))(*layer1)
            check(self.lib.hg_gemm_plan_create_step(a1, a2, a3, harr, len(l2), C.byref(handle)))
        else:
            check(self.lib.hg_gemm_plan_create_tail(a2, a3, harr, len(l2), C.byref(handle)))
        self._all_plans.append(handle.value)
        return handle.value

    def run_head_plan(self, plan: int, eps, seed: int, counter: int) -> None:
        check(self.lib.hg_gemm_plan_run_heads(plan, ptr(eps), seed, counter, current_stream()))

    def run_plans(self, plans: List[int]) -> None:
        st = current_stream()
        run = self.lib.hg_gemm_plan_run
        for h in plans:
            if run(h, st) != 0:
                check(-1)

    def __del__(self):
        try:
            for h in getattr(self, "_all_plans", []):
                self.lib.hg_gemm_plan_destroy(h)
        except Exception:
            pass

    # -- forward of all hidden layers ---------------------------------------------------------
    def forward_hidden(self, nets: Sequence[NetSpec], inputs: Dict[str, torch.Tensor], in_cols: Dict[str, int],
                       ws: Workspace, rows: int, make_plans: bool = False, heads=None, frames: Optional[Dict[str, tuple]] = None):
        """inputs[name]: bf16 [rows, ld] (ld multiple of 8; only the first in_cols[name] columns are read).
        Fills ws.h[name][l] = elu(W_l h_{l-1} + b_l) for every hidden layer, one grouped launch per depth.
        With make_plans=True nothing is launched: the prepared launches are returned for `run_plans`.
        heads (plans only): {net name: HgFusedHead} -- the LAST depth becomes one fused launch (last hidden layer +
        output layer + sampling, `run_head_plan`); its activations are not written to ws.h.
        frames: {net name: (width, pitch)} -- the input rows of that net are history-ring frames of `width` columns at a
        pitch of `pitch` (in_cols = frames * pitch): layer 1 reads the frame-pitched weight shadow."""
        plans = []
        depth = max(n.num_hidden for n in nets)
        fuse_tail = (make_plans and heads is not None and self.can_fuse_tail(nets))
        for l in range(depth):
            if fuse_tail and l == depth - 2:
                plans.append(self._tail_plan(nets, ws, rows, heads))     # layers depth-1, depth and the heads: one launch
                break
            probs, ns = [], []
            for net in nets:
                if l >= net.num_hidden:
                    continue
                lin = net.linears[l]
                a = inputs[net.name] if l == 0 else ws.h[net.name][l - 1]
                k = in_cols[net.name] if l == 0 else net.hidden[l - 1]
                fr = frames.get(net.name) if (frames and l == 0) else None
                w16 = self.flat.framed_shadow(lin.weight, *fr) if fr else self.flat.w16[id(lin.weight)]
                out = ws.h[net.name][l]
                probs.append(HgGemmProblem(
                    A=ptr(a), B=ptr(w16), lda=a.stride(0), ldb=w16.stride(0), a_mn_major=0, b_mn_major=0,
                    M=rows, N=lin.out_features, K=k, k_splits=1, out=ptr(out), ldo=out.stride(0),
                    bias=ptr(lin.bias.data), act=0, ldact=0, colsum=0, out_frame_src=0, out_frame_dst=0))
                ns.append(lin.out_features)
            if fuse_tail and depth == 3 and l == 0 and os.environ.get("HG_ROLLOUT_STEP", "0") == "1":
                # layer 1 + tail in ONE launch (rollout_step_kernel); shapes that do not qualify keep the two launches.
                # OFF by default: bit-identical to the two launches, but measured SLOWER at 4096 envs (rollout 2.90 vs 2.39 ms
                # per iteration): layer 1 is bound by the L2 -> SM operand fill per SM and takes 20 us on the 84 SMs the 64 tail
                # CTAs leave it instead of 13.5 us on 148; the overlap it buys the tail is smaller than that (device timeline)
                try:
                    plans.append(self._tail_plan(nets, ws, rows, heads, layer1=probs))
                    break
                except _lib.HgError:
                    pass
            if make_plans and heads is not None and l == depth - 1:
                hs = [heads[net.name] for net in nets if l < net.num_hidden]
                arr, harr = (HgGemmProblem * len(probs))(*probs), (_lib.HgFusedHead * len(hs))(*hs)
                handle = C.c_void_p()
                check(self.lib.hg_gemm_plan_create_heads(arr, harr, len(probs), C.byref(handle)))
                self._all_plans.append(handle.value)
                plans.append(handle.value)
            elif make_plans:
                plans.append(self._plan(probs, EPI_BIAS_ELU_BF16, self._block_n(ns, "fwd", rows)))
            else:
                self._launch(probs, EPI_BIAS_ELU_BF16, self._block_n(ns, "fwd", rows))
        return plans

    # -- backward of all hidden layers ----------------------------------------------------------
    def backward_hidden(self, nets: Sequence[NetSpec], inputs: Dict[str, torch.Tensor], in_cols: Dict[str, int],
                        ws: Workspace, rows: int, frames: Optional[Dict[str, tuple]] = None) -> None:
        """Expects ws.dz[name][-1] (top hidden layer pre-activation gradient) and its bias gradient to be
        already produced by the head kernel.  Accumulates weight gradients into the gradient arena and
        bias gradients of the lower layers, walking down one layer per step, grouped across nets."""
        depth = max(n.num_hidden for n in nets)
        for s in range(depth):
            dw, dx, dw_n, dx_n = [], [], [], []
            for net in nets:
                j = net.num_hidden - 1 - s            # hidden layer index whose dZ is available
                if j < 0:
                    continue
                lin = net.linears[j]
                dz = ws.dz[net.name][j]
                below = inputs[net.name] if j == 0 else ws.h[net.name][j - 1]
                k_in = in_cols[net.name] if j == 0 else net.hidden[j - 1]
                gw = self.flat.grad_view(lin.weight)
                fr = frames.get(net.name) if (frames and j == 0) else None
                # dW_j [out, in] += dZ_j^T [out, rows] * H_{j-1} [rows, in]     (both operands MN-major); frame-pitched input
                # columns (history ring) are mapped back to the packed columns of the weight gradient by the epilogue
                dw.append(HgGemmProblem(
                    A=ptr(dz), B=ptr(below), lda=dz.stride(0), ldb=below.stride(0), a_mn_major=1, b_mn_major=1,
                    M=lin.out_features, N=k_in, K=rows, k_splits=1, out=ptr(gw), ldo=lin.in_features,
                    bias=0, act=0, ldact=0, colsum=0, out_frame_src=fr[1] if fr else 0, out_frame_dst=fr[0] if fr else 0))
                dw_n.append(k_in)
                if j >= 1:
                    # dZ_{j-1} = (dZ_j [rows,out] * W_j [out,in]) . elu'(H_{j-1});  b-grad of layer j-1 = colsum
                    w16 = self.flat.w16[id(lin.weight)]
                    out = ws.dz[net.name][j - 1]
                    gb = self.flat.grad_view(net.linears[j - 1].bias)
                    dx.append(HgGemmProblem(
                        A=ptr(dz), B=ptr(w16), lda=dz.stride(0), ldb=w16.stride(0), a_mn_major=0, b_mn_major=1,
                        M=rows, N=k_in, K=lin.out_features, k_splits=1, out=ptr(out), ldo=out.stride(0),
                        bias=0, act=ptr(below), ldact=below.stride(0), colsum=ptr(gb), out_frame_src=0, out_frame_dst=0))
                    dx_n.append(k_in)
            if dw:
                bn = self._block_n(dw_n, "dw")
                cg = self._cta_group(dw, EPI_ATOMIC_F32)
                tiles = sum(((p.M + 128 * cg - 1) // (128 * cg)) * ((p.N + bn - 1) // bn) for p in dw)
                # split-K so that every CTA (pair) gets `waves` tiles.  One tile per CTA leaves the fp32-atomic epilogue of
                # the tile exposed (9 us of a 31 us layer-3 launch in the device timeline); two or more would overlap it
                # with the next tile's main loop, but measured SLOWER (update 4.84 / 4.95 / 5.00 / 5.14 ms for 1 / 2 / 3 / 4
                # waves: every extra split adds a full fp32 RED pass over the weight gradient), so 1 stays
                waves = int(os.environ.get("HG_DW_WAVES", "1"))
                splits = max(1, min((rows + 63) // 64, waves * (self.num_sms // cg) // max(tiles, 1)))
                for p in dw:
                    p.k_splits = splits
                self._launch(dw, EPI_ATOMIC_F32, bn)
            if dx:
                self._launch(dx, EPI_DELU_BF16, self._block_n(dx_n, "dx"))
