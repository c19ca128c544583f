"""Projected Levenberg-Marquardt solve of the FTE problem on the GPU(s) - host driver.

Replaces the reference's `opt.solve(m)` (Pyomo -> IPOPT subprocess,
/root/reference/src/all_optimizations.py:503-524) for the reduced objective

    F(x) = sum_{n,c,l,d} rho(w (proj - meas))  +  sum_{n>=3,p} (1/Q_p) (third difference / Ts^2)^2

(measurement term :394-399,494-497; dynamics :369-391; weights :245-252,310-315) under the 21 pose
bounds (:403-483).  Per attempt: lm_assemble -> block cyclic reduction (csrc/bcr.cu) -> lm_step ->
fte_eval at the trial point -> lm_prepare -> lm_reduce; the host only reads 5 scalars back.

Multi-GPU: frames are sharded in contiguous blocks (multiples of 3 frames); every rank reduces its
shard onto its two end super-blocks (pinned BCR), ONE all_gather moves the interface blocks
(2 x (75x75 + 75x75 + 75) doubles per rank), every rank solves the 2G-block interface chain
redundantly (bit-identical, no broadcast needed) and back-substitutes locally.  The interface
solution of the neighbouring rank is exactly the step of the halo frames, so no separate halo
exchange is needed.  A second, 5-scalar all_reduce carries the objective / model-reduction sums.
"""
import numpy as np

from . import bcr as _bcr
from . import _lib

NA = _lib.N_ACTIVE
SB = _bcr.SB

# measurement-model std-devs (all_optimizations.py:245-252), active slots only; weight = 1/sigma^2
Q_SIGMA_ACTIVE = np.array([4, 7, 5, 13, 32, 10, 9, 18, 43, 53, 90, 118, 247, 186, 194, 164, 295, 243, 334, 149,
                           26, 12, 34, 43, 51], dtype=np.float64)
_P6, _P15, _P2, _PI = np.pi / 6, np.pi / 1.5, np.pi / 2, np.pi


def default_bounds():
    """(lo, hi)[25] of all_optimizations.py:403-483 in the active ordering; +-inf where free."""
    lo = np.full(NA, -np.inf)
    hi = np.full(NA, np.inf)
    for i in (3, 4, 5, 6, 7, 8, 9, 21, 22):          # phi0 phi1 phi3 theta0..3 psi1 psi3
        lo[i], hi[i] = -_P6, _P6
    for i in (10, 11, 23, 24):                        # theta4 theta5 psi4 psi5
        lo[i], hi[i] = -_P15, _P15
    for i in (12, 14, 16, 18):                        # shoulders / hips
        lo[i], hi[i] = -_P2, _P2
    for i in (13, 15):                                # front knees
        lo[i], hi[i] = -_PI, 0.0
    for i in (17, 19):                                # back knees
        lo[i], hi[i] = 0.0, _PI
    return lo, hi


def shard_frames(n_global, world):
    """Contiguous shards in multiples of 3 frames (super-block aligned); the last rank takes the rest.
    Returns a list of (frame0, n_frames)."""
    per = -(-n_global // world)          # ceil
    per = -(-per // 3) * 3
    out = []
    start = 0
    for r in range(world):
        n = min(per, max(n_global - start, 0)) if r < world - 1 else max(n_global - start, 0)
        out.append((start, n))
        start += n
    return out


# ---- interface chain plumbing (pure torch: runs on CPU tensors with gloo in the tests) ----------
PAYLOAD = 4 * SB * SB + 4 * SB     # D_f, D_l, Lc_f, Lc_l, rhs_f, rhs_l, fixed_f, fixed_l


def pack_interface(D, Lc, rhs, fixed_blocks):
    """Local reduced system (end super-blocks 0 and M-1) -> flat payload tensor."""
    import torch

    M = D.shape[0]
    return torch.cat([D[0].reshape(-1), D[M - 1].reshape(-1), Lc[0].reshape(-1), Lc[M - 1].reshape(-1),
                      rhs[0], rhs[M - 1], fixed_blocks[0].to(D.dtype), fixed_blocks[M - 1].to(D.dtype)])


def gather_interface_chain(payload, world, group=None):
    """all_gather the payloads and build the 2G-block interface chain (D, Lc, rhs)."""
    import torch
    import torch.distributed as dist

    allp = torch.empty(world * PAYLOAD, dtype=payload.dtype, device=payload.device)
    if world > 1:
        dist.all_gather_into_tensor(allp, payload.contiguous(), group=group)
    else:
        allp.copy_(payload)
    allp = allp.view(world, PAYLOAD)
    n2 = SB * SB
    D = torch.empty(2 * world, SB, SB, dtype=payload.dtype, device=payload.device)
    Lc = torch.zeros_like(D)
    rhs = torch.empty(2 * world, SB, dtype=payload.dtype, device=payload.device)
    D[0::2] = allp[:, 0:n2].view(world, SB, SB)
    D[1::2] = allp[:, n2:2 * n2].view(world, SB, SB)
    Lc[0::2] = allp[:, 2 * n2:3 * n2].view(world, SB, SB)
    Lc[1::2] = allp[:, 3 * n2:4 * n2].view(world, SB, SB)
    rhs[0::2] = allp[:, 4 * n2:4 * n2 + SB]
    rhs[1::2] = allp[:, 4 * n2 + SB:4 * n2 + 2 * SB]
    fixed_l = allp[:, 4 * n2 + 3 * SB:4 * n2 + 4 * SB]
    # the coupling of rank r's first block to rank r-1's last block: zero the columns of the
    # variables rank r-1 froze (rank r could not know them when it assembled the block)
    if world > 1:
        keep = (1.0 - fixed_l[:-1]).unsqueeze(1)          # (G-1, 1, 75)
        Lc[2::2] = Lc[2::2] * keep
    Lc[0].zero_()
    return D, Lc, rhs


def split_interface_solution(x_chain, rank, world):
    """-> (x_first, x_last, halo_left or None, halo_right or None) for this rank."""
    xf, xl = x_chain[2 * rank], x_chain[2 * rank + 1]
    hl = x_chain[2 * rank - 1] if rank > 0 else None
    hr = x_chain[2 * rank + 2] if rank < world - 1 else None
    return xf, xl, hl, hr


class ChainSolver:
    """BCR solve of a block-tridiagonal chain held in device tensors (CUDA kernels of csrc/bcr.cu)."""

    def __init__(self, handle, M, pinned=False):
        import torch

        self.h = handle
        self.M = M
        self.pinned = pinned
        dev = torch.device("cuda", handle.device)
        levels, left = _bcr.make_schedule(M, pin_first=pinned, pin_last=pinned)
        self.levels = [dict(elim=torch.from_numpy(lv["elim"]).to(dev), surv=torch.from_numpy(lv["surv"]).to(dev))
                       for lv in levels]
        self.left = left
        self.P = torch.zeros(M, SB, SB, dtype=torch.float64, device=dev)
        self.Q = torch.zeros(M, SB, SB, dtype=torch.float64, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)

    def reduce(self, D, Lc, rhs):
        for lv in self.levels:
            ne, ns = lv["elim"].shape[0], lv["surv"].shape[0]
            self.h.call_dev("acino_bcr_factor_dev", ne, lv["elim"], D, Lc, self.P, self.Q, rhs, self.info)
            if ns:
                self.h.call_dev("acino_bcr_update_dev", ns, lv["surv"], D, Lc, self.P, self.Q, rhs)

    def backsub(self, D, rhs, x):
        for lv in reversed(self.levels):
            self.h.call_dev("acino_bcr_backsub_dev", lv["elim"].shape[0], lv["elim"], D, self.P, self.Q, rhs, x)

    def solve(self, D, Lc, rhs, x):
        """Unpinned chain: full solve (D, Lc, rhs are overwritten)."""
        self.reduce(D, Lc, rhs)
        self.backsub(D, rhs, x)


class FTESolver:
    """LM solver for one shard of frames on one GPU (world = 1: the whole problem)."""

    def __init__(self, handle, meas, w, Ts, frame0=0, n_global=None, q=None, bounds=None, rank=0, world=1,
                 group=None):
        import torch

        self.h = handle
        self.torch = torch
        self.dev = torch.device("cuda", handle.device)
        self.N = int(meas.shape[0])
        self.frame0 = int(frame0)
        self.ng = int(n_global if n_global is not None else self.N)
        self.rank, self.world, self.group = rank, world, group
        if world > 1 and (self.N % 3 != 0 and rank < world - 1):
            raise ValueError("shards must be multiples of 3 frames (use shard_frames)")
        if world > 1 and self.N < 6:
            raise ValueError("need at least 6 frames per rank")
        self.M = -(-self.N // 3)
        f64, f32 = torch.float64, torch.float32
        dev = self.dev
        self.meas = torch.as_tensor(np.ascontiguousarray(meas, dtype=np.float32)).to(dev)
        self.w = torch.as_tensor(np.ascontiguousarray(w, dtype=np.float32)).to(dev)
        q = 1.0 / Q_SIGMA_ACTIVE ** 2 if q is None else np.asarray(q, dtype=np.float64)
        lo, hi = default_bounds() if bounds is None else bounds
        self.sw = torch.as_tensor(2.0 * q / Ts ** 4, dtype=f64).to(dev)
        self.lo = torch.as_tensor(np.asarray(lo, dtype=np.float64)).to(dev)
        self.hi = torch.as_tensor(np.asarray(hi, dtype=np.float64)).to(dev)
        N, M = self.N, self.M

        def buf(*shape, dtype=f64):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        # two states (accepted / trial), swapped on acceptance
        self.st = [dict(x_ext=buf(N + 6, NA), x32=buf(N, NA, dtype=f32), cost=buf(N, dtype=f32),
                        g=buf(N, NA, dtype=f32), H=buf(N, _lib.N_UPPER, dtype=f32), gtot=buf(N, NA),
                        fixed=buf(N, NA, dtype=torch.uint8), cost_s=buf(N)) for _ in range(2)]
        self.d_ext = buf(N + 6, NA)
        self.pred, self.step = buf(N), buf(N)
        self.D, self.Lc = buf(M, SB, SB), buf(M, SB, SB)
        self.rhs, self.dx = buf(M, SB), buf(M, SB)
        self.out5 = buf(5)
        self.local = ChainSolver(handle, M, pinned=world > 1)
        self.iface = ChainSolver(handle, 2 * world, pinned=False) if world > 1 else None
        self.n_launch0 = handle.launch_count

    # -- pieces ---------------------------------------------------------------------------------
    def _eval(self, s):
        self.h.fte_eval_dev(s["x32"], self.meas, self.w, s["cost"], s["g"], s["H"])

    def _prepare(self, s):
        self.h.call_dev("acino_lm_prepare_dev", self.N, self.frame0, self.ng, s["x_ext"], s["g"], self.sw, self.lo,
                        self.hi, s["gtot"], s["fixed"], s["cost_s"])

    def _sums(self, s, with_step):
        """Global (F, pred, step_inf): fixed-order local sums + one 5-scalar all_reduce."""
        torch = self.torch
        self.h.call_dev("acino_lm_reduce_dev", self.N, s["cost"], s["cost_s"], self.pred if with_step else None, None,
                        self.step if with_step else None, self.out5)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(self.out5[:4], op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.out5[4:], op=dist.ReduceOp.MAX, group=self.group)
        o = self.out5.cpu().numpy()
        return float(o[0] + o[1]), float(o[2]), float(o[4])

    def _solve_step(self, s, lam):
        """(B + lam diag B) dx = -g on the accepted state -> self.d_ext (with halos)."""
        torch = self.torch
        N, M = self.N, self.M
        self.h.call_dev("acino_lm_assemble_dev", N, self.frame0, self.ng, M, s["H"], s["gtot"], s["fixed"], self.sw,
                        float(lam), self.D, self.Lc, self.rhs)
        self.d_ext.zero_()
        if self.world == 1:
            self.local.solve(self.D, self.Lc, self.rhs, self.dx)
        else:
            self.local.reduce(self.D, self.Lc, self.rhs)
            fixed_blocks = torch.zeros(M * 3, NA, dtype=torch.uint8, device=self.dev)
            fixed_blocks[:N] = s["fixed"]
            payload = pack_interface(self.D, self.Lc, self.rhs, fixed_blocks.view(M, SB))
            Dc, Lcc, rc = gather_interface_chain(payload, self.world, self.group)
            xc = torch.zeros(2 * self.world, SB, dtype=torch.float64, device=self.dev)
            self.iface.solve(Dc, Lcc, rc, xc)
            xf, xl, hl, hr = split_interface_solution(xc, self.rank, self.world)
            self.dx[0], self.dx[M - 1] = xf, xl
            self.local.backsub(self.D, self.rhs, self.dx)
            if hl is not None:
                self.d_ext[0:3] = hl.view(3, NA)
            if hr is not None:
                self.d_ext[N + 3:N + 6] = hr.view(3, NA)
        self.d_ext[3:3 + N] = self.dx.view(-1, NA)[:N]

    def _trial(self, s, t):
        torch = self.torch
        N = self.N
        self.h.call_dev("acino_lm_step_dev", N, self.frame0, self.ng, s["x_ext"], self.d_ext, s["gtot"], s["H"], self.sw,
                        self.lo, self.hi, t["x_ext"], t["x32"], self.pred, self.step)
        # halo rows of the trial state: the neighbour applies the same clamp to the same numbers
        t["x_ext"][0:3] = torch.minimum(torch.maximum(s["x_ext"][0:3] + self.d_ext[0:3], self.lo), self.hi)
        t["x_ext"][N + 3:] = torch.minimum(torch.maximum(s["x_ext"][N + 3:] + self.d_ext[N + 3:], self.lo), self.hi)
        self._eval(t)
        self._prepare(t)

    def _exchange_halo_init(self, s, x0):
        """One-time halo fill of the initial iterate (all_gather of the 3 first / last frames)."""
        torch = self.torch
        if self.world == 1:
            return
        import torch.distributed as dist

        N = self.N
        mine = torch.cat([s["x_ext"][3:6].reshape(-1), s["x_ext"][N:N + 3].reshape(-1)])
        allb = torch.empty(self.world * mine.numel(), dtype=mine.dtype, device=self.dev)
        dist.all_gather_into_tensor(allb, mine, group=self.group)
        allb = allb.view(self.world, 2, 3, NA)
        if self.rank > 0:
            s["x_ext"][0:3] = allb[self.rank - 1, 1]
        if self.rank < self.world - 1:
            s["x_ext"][N + 3:N + 6] = allb[self.rank + 1, 0]

    # -- the loop -------------------------------------------------------------------------------
    def solve(self, x0, max_iter=60, lam0=1e-3, tol_step=1e-6, tol_rel=1e-8, max_attempts=12, verbose=False):
        torch = self.torch
        N = self.N
        s, t = self.st
        x0 = np.clip(np.asarray(x0, dtype=np.float64), self.lo.cpu().numpy(), self.hi.cpu().numpy())
        s["x_ext"].zero_()
        s["x_ext"][3:3 + N] = torch.as_tensor(x0).to(self.dev)
        s["x32"].copy_(s["x_ext"][3:3 + N].to(torch.float32))
        self._exchange_halo_init(s, x0)
        self._eval(s)
        self._prepare(s)
        F, _, _ = self._sums(s, with_step=False)
        lam = lam0
        hist = [F]
        n_eval, n_solve, it = 1, 0, 0
        converged = False
        for it in range(max_iter):
            accepted = False
            for _ in range(max_attempts):
                self._solve_step(s, lam)
                n_solve += 1
                self._trial(s, t)
                n_eval += 1
                Ft, pred, step = self._sums(t, with_step=True)
                rho = (F - Ft) / pred if pred > 0 else -1.0
                if verbose and self.rank == 0:
                    print(f"it {it:3d} lam {lam:9.3e} F {F:16.6f} Ft {Ft:16.6f} pred {pred:10.3e} rho {rho:7.3f} |dx|inf {step:.2e}")
                if Ft < F and rho > 1e-4:
                    accepted = True
                    rel = (F - Ft) / max(abs(F), 1e-30)
                    F = Ft
                    s, t = t, s
                    if rho > 0.75:
                        lam = max(lam / 3, 1e-12)
                    elif rho < 0.25:
                        lam *= 2
                    break
                lam *= 4
            hist.append(F)
            if not accepted:
                break
            if step < tol_step or rel < tol_rel:
                converged = True
                break
        self.st = [s, t]
        torch.cuda.synchronize(self.dev)
        x = s["x_ext"][3:3 + N].cpu().numpy()
        info = dict(F=F, iters=it + 1, n_eval=n_eval, n_solve=n_solve, history=hist, lam=lam, converged=converged,
                    bcr_info=int(self.local.info.item()), launches=self.h.launch_count - self.n_launch0)
        return x, info
