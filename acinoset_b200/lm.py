"""Projected Levenberg-Marquardt solve of the FTE problem on the GPU(s) - host driver.

Replaces the reference's `opt.solve(m)` (Pyomo -> IPOPT subprocess,
/root/reference/src/all_optimizations.py:503-524) for the reduced objective

    F(x) = sum_{n,c,l,d} rho(w (proj - meas))  +  sum_{n>=3,p} (1/Q_p) (third difference / Ts^2)^2

(measurement term :394-399,494-497; dynamics :369-391; weights :245-252,310-315) under the 21 pose
bounds (:403-483).

One attempt = the phases of csrc/lm_plan.cu, all enqueued on one stream with no host decision in
between: REDUCE (structured level 0 fused with the assembly, csrc/lm_l0.cu, then the dense block
cyclic reduction levels, csrc/bcr.cu) -> BACKSUB -> TRIAL (projected step, fte_eval at the trial
point, gradient / frozen set, fixed-order sums) -> DECIDE (gain ratio, lambda update, commit,
convergence test - on the device).  After one eager attempt the sequence is captured into a CUDA
graph and replayed; the host stays one attempt ahead of the GPU and only looks at the pinned mirror
of the device control block to learn that the solve has finished.

Multi-GPU: frames are sharded in contiguous blocks (multiples of 3 frames); every rank reduces its
shard onto its two end super-blocks (pinned BCR), one all_gather moves the interface blocks
(22 800 doubles per rank), every rank solves the 2G-block interface chain redundantly
(bit-identical, no broadcast needed) and back-substitutes locally; the interface solution of the
neighbouring rank is exactly the step of the halo frames.  A second all_gather of 8 doubles per
rank carries the objective / model-reduction partials of the trial point: the accept / reject
decision of attempt k needs the evaluation that FOLLOWS the interface exchange of attempt k, and
the assembly of attempt k+1 needs that decision, so two exchange points per attempt are inherent
(both are captured in the graph; every rank sums the partials in rank order => identical bits).
"""
import ctypes
import numpy as np

from . import bcr as _bcr
from . import _lib

NA = _lib.N_ACTIVE
SB = _bcr.SB

from .config import CHEETAH

# model std-devs (all_optimizations.py:245-252), active slots only; weight = 1/sigma^2 (one definition: config.py)
Q_SIGMA_ACTIVE = CHEETAH.q_sigma


def default_bounds():
    """(lo, hi)[25] of all_optimizations.py:403-483 in the active ordering; +-inf where free."""
    lo, hi = CHEETAH.bounds
    return lo.copy(), hi.copy()


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
        self.R = torch.zeros(M, SB, SB, dtype=torch.float64, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)

    def reduce(self, D, Lc, rhs):
        for lv in self.levels:
            ne, ns = lv["elim"].shape[0], lv["surv"].shape[0]
            self.h.call_dev("acino_bcr_factor_dev", ne, lv["elim"], D, Lc, self.P, self.Q, self.R, rhs, self.info)
            if ns:
                self.h.call_dev("acino_bcr_update_dev", ns, lv["surv"], D, Lc, self.P, self.Q, rhs)

    def backsub(self, D, rhs, x):
        for lv in reversed(self.levels):
            self.h.call_dev("acino_bcr_backsub_dev", lv["elim"].shape[0], lv["elim"], self.R, self.P, self.Q, rhs, x)

    def solve(self, D, Lc, rhs, x):
        """Unpinned chain: full solve (D, Lc, rhs are overwritten)."""
        self.reduce(D, Lc, rhs)
        self.backsub(D, rhs, x)




# ---- device control block (csrc/lm_common.cuh enum LmCtl) and phases (include/acino_b200.h) ------
(CTL_LAM, CTL_F, CTL_FT, CTL_PRED, CTL_STEP, CTL_RHO, CTL_REL, CTL_ACCEPT, CTL_DONE, CTL_N_ATTEMPT, CTL_N_ACCEPT,
 CTL_FAIL_STREAK, CTL_STATUS, CTL_MAX_ITER, CTL_MAX_ATTEMPTS, CTL_TOL_STEP, CTL_TOL_REL, CTL_ITERS, CTL_HIST_CAP,
 CTL_TOL_NOISE, CTL_NOISE_STREAK, CTL_N_ENQ, CTL_DONE_AT) = range(23)
CTL_SIZE, N_SUMS, N_HIST = 32, 8, 8
PH_INIT_EVAL, PH_INIT_FINISH, PH_REDUCE, PH_BACKSUB, PH_TRIAL, PH_DECIDE = range(6)
STATUS = {0: "running", 1: "converged", 2: "no acceptable step", 3: "iteration limit"}


def level0_split(M, pinned):
    """(elim0, surv0, dense levels): the first elimination level is done by the structured kernels of
    csrc/lm_l0.cu, which also ASSEMBLE every block that survives it - so surv0 lists every block that is not
    eliminated at level 0 with its eliminated neighbours (or -1), not only the ones that receive an update."""
    levels, left = _bcr.make_schedule(M, pin_first=pinned, pin_last=pinned)
    elim0 = levels[0]["elim"] if levels else np.zeros((0, 3), np.int32)
    gone = set(int(e) for e in elim0[:, 0])
    surv0 = np.array([(j, j - 1 if (j - 1) in gone else -1, j + 1 if (j + 1) in gone else -1)
                      for j in range(M) if j not in gone], dtype=np.int32).reshape(-1, 3)
    return elim0, surv0, levels[1:], left


def flatten_levels(levels):
    """-> (counts (L,2) int32 host array, sched (sum(ne+ns),3) int32: per level the elim rows, then the surv rows)."""
    counts = np.array([(lv["elim"].shape[0], lv["surv"].shape[0]) for lv in levels], dtype=np.int32).reshape(-1, 2)
    rows = [r for lv in levels for r in (lv["elim"], lv["surv"])]
    sched = np.concatenate(rows).astype(np.int32) if rows else np.zeros((0, 3), np.int32)
    return np.ascontiguousarray(counts), np.ascontiguousarray(sched.reshape(-1, 3))


_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64


class LmDesc(ctypes.Structure):
    """acino_lm_desc (include/acino_b200.h)."""
    _fields_ = ([("n_frames", _i32), ("n_blocks", _i32), ("rank", _i32), ("world", _i32), ("frame0", _i64), ("n_global", _i64)]
                + [(k, _vp) for k in ("meas", "w", "sw", "lo", "hi")]
                + [(k, _vp * 2) for k in ("x_ext", "x32", "cost", "g", "H", "gtot", "fixed", "cost_s")]
                + [(k, _vp) for k in ("pred", "step", "D", "Lc", "P", "Q", "R", "rhs", "dx", "dhalo", "info")]
                + [("n_elim0", _i32), ("n_surv0", _i32), ("elim0", _vp), ("surv0", _vp),
                   ("n_levels", _i32), ("level_counts", _vp), ("sched", _vp)]
                + [(k, _vp) for k in ("payload", "gathered", "cD", "cLc", "cP", "cQ", "cR", "crhs", "cx")]
                + [("n_clevels", _i32), ("clevel_counts", _vp), ("csched", _vp)]
                + [(k, _vp) for k in ("sums_local", "sums_all", "ctl", "ctl_host", "hist")]
                + [("hist_cap", _i32)])


_lib.lib.acino_lm_desc_size.argtypes = []
_lib.lib.acino_lm_desc_size.restype = ctypes.c_int
if _lib.lib.acino_lm_desc_size() != ctypes.sizeof(LmDesc):
    raise ImportError("acino_lm_desc layout mismatch between include/acino_b200.h and acinoset_b200/lm.py")
_lib.lib.acino_lm_plan_create.argtypes = [ctypes.c_void_p, ctypes.POINTER(LmDesc), ctypes.POINTER(ctypes.c_void_p)]
_lib.lib.acino_lm_plan_create.restype = ctypes.c_int
_lib.lib.acino_lm_plan_destroy.argtypes = [ctypes.c_void_p]
_lib.lib.acino_lm_plan_destroy.restype = ctypes.c_int
_lib.lib.acino_lm_enqueue.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
_lib.lib.acino_lm_enqueue.restype = ctypes.c_int


class FTESolver:
    """LM solver for one shard of frames on one GPU (world = 1: the whole problem)."""

    HIST_CAP = 4096

    def __init__(self, handle, meas, w, Ts, frame0=0, n_global=None, q=None, bounds=None, rank=0, world=1,
                 group=None, use_graph=True):
        import torch

        self.h = handle
        self._scene_version = getattr(handle, "scene_version", 0)   # the captured graph holds THIS scene by value
        self.torch = torch
        self.dev = torch.device("cuda", handle.device)
        self.N = int(meas.shape[0])
        self.frame0 = int(frame0)
        self.ng = int(n_global if n_global is not None else self.N)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.use_graph = bool(use_graph)
        if world > 1 and (self.N % 3 != 0 and rank < world - 1):
            raise ValueError("shards must be multiples of 3 frames (use shard_frames)")
        if world > 1 and self.N < 6:
            raise ValueError("need at least 6 frames per rank")
        if self.N < 1:
            raise ValueError("need at least one frame")
        self.M = -(-self.N // 3)
        f64, f32, i32 = torch.float64, torch.float32, torch.int32
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

        # two states: [0] accepted, [1] trial (DECIDE copies trial -> accepted on the device when a step is accepted)
        self.st = [dict(x_ext=buf(N + 6, NA), x32=buf(N, NA, dtype=f32), cost=buf(N, dtype=f32),
                        g=buf(N, NA, dtype=f32), H=buf(N, _lib.N_UPPER, dtype=f32), gtot=buf(N, NA),
                        fixed=buf(N, NA, dtype=torch.uint8), cost_s=buf(N)) for _ in range(2)]
        self.pred, self.step = buf(N), buf(N)
        self.D, self.Lc, self.P, self.Q, self.R = (buf(M, SB, SB) for _ in range(5))
        self.rhs, self.dx = buf(M, SB), buf(M, SB)
        self.dhalo = buf(2, SB)
        self.info = torch.zeros(1, dtype=i32, device=dev)
        self.sums_local = buf(N_SUMS)
        self.sums_all = buf(world, N_SUMS) if world > 1 else self.sums_local
        self.ctl = buf(CTL_SIZE)
        self.ctl_host = torch.zeros(CTL_SIZE, dtype=f64).pin_memory()
        self.hist = buf(self.HIST_CAP, N_HIST)
        # schedules
        elim0, surv0, levels, _ = level0_split(M, pinned=world > 1)
        self._counts, sched = flatten_levels(levels)
        self.elim0 = torch.from_numpy(np.ascontiguousarray(elim0)).to(dev)
        self.surv0 = torch.from_numpy(np.ascontiguousarray(surv0)).to(dev)
        self.sched = torch.from_numpy(sched).to(dev)
        self.n_launch0 = handle.launch_count
        d = LmDesc()
        d.n_frames, d.n_blocks, d.rank, d.world, d.frame0, d.n_global = N, M, self.rank, self.world, self.frame0, self.ng
        for k, t in (("meas", self.meas), ("w", self.w), ("sw", self.sw), ("lo", self.lo), ("hi", self.hi),
                     ("pred", self.pred), ("step", self.step), ("D", self.D), ("Lc", self.Lc), ("P", self.P), ("Q", self.Q), ("R", self.R),
                     ("rhs", self.rhs), ("dx", self.dx), ("dhalo", self.dhalo), ("info", self.info),
                     ("elim0", self.elim0), ("surv0", self.surv0), ("sched", self.sched),
                     ("sums_local", self.sums_local), ("sums_all", self.sums_all), ("ctl", self.ctl),
                     ("ctl_host", self.ctl_host), ("hist", self.hist)):
            setattr(d, k, t.data_ptr())
        for k in ("x_ext", "x32", "cost", "g", "H", "gtot", "fixed", "cost_s"):
            setattr(d, k, (ctypes.c_void_p * 2)(self.st[0][k].data_ptr(), self.st[1][k].data_ptr()))
        d.n_elim0, d.n_surv0, d.n_levels = int(elim0.shape[0]), int(surv0.shape[0]), int(self._counts.shape[0])
        d.level_counts = self._counts.ctypes.data
        d.hist_cap = self.HIST_CAP
        if world > 1:
            self.payload = buf(PAYLOAD)
            self.gathered = buf(world, PAYLOAD)
            G2 = 2 * world
            self.cD, self.cLc, self.cP, self.cQ, self.cR = (buf(G2, SB, SB) for _ in range(5))
            self.crhs, self.cx = buf(G2, SB), buf(G2, SB)
            clevels, _ = _bcr.make_schedule(G2)
            self._ccounts, csched = flatten_levels(clevels)
            self.csched = torch.from_numpy(csched).to(dev)
            for k in ("payload", "gathered", "cD", "cLc", "cP", "cQ", "cR", "crhs", "cx", "csched"):
                setattr(d, k, getattr(self, k).data_ptr())
            d.n_clevels = int(self._ccounts.shape[0])
            d.clevel_counts = self._ccounts.ctypes.data
        self._plan = ctypes.c_void_p()
        handle._check(_lib.lib.acino_lm_plan_create(handle._h, ctypes.byref(d), ctypes.byref(self._plan)),
                      "acino_lm_plan_create")
        self._graph = None

    def close(self):
        """Release the captured CUDA graph and the plan.  With world > 1 the graph holds NCCL work: close every solver
        BEFORE torch.distributed.destroy_process_group() (destroying a communicator that a live graph references blocks)."""
        self._graph = None
        if getattr(self, "_plan", None) is not None and self._plan.value:
            _lib.lib.acino_lm_plan_destroy(self._plan)
            self._plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- pieces (also used by the tests) ------------------------------------------------------------
    def _eval(self, s):
        self.h.fte_eval_dev(s["x32"], self.meas, self.w, s["cost"], s["g"], s["H"])

    def _prepare(self, s):
        self.h.call_dev("acino_lm_prepare_dev", self.N, self.frame0, self.ng, s["x_ext"], s["g"], self.sw, self.lo,
                        self.hi, s["gtot"], s["fixed"], s["cost_s"])

    def _enq(self, phase):
        st = self.torch.cuda.current_stream(self.dev).cuda_stream
        self.h._check(_lib.lib.acino_lm_enqueue(self.h._h, self._plan, int(phase), ctypes.c_void_p(st)), "acino_lm_enqueue")

    def _gather(self, out, inp):
        if self.world > 1:
            import torch.distributed as dist

            dist.all_gather_into_tensor(out.view(-1), inp.view(-1), group=self.group)

    def _attempt(self):
        """Enqueue one LM attempt (no host decision anywhere in it)."""
        self._enq(PH_REDUCE)
        if self.world > 1:
            self._gather(self.gathered, self.payload)
        self._enq(PH_BACKSUB)
        self._enq(PH_TRIAL)
        if self.world > 1:
            self._gather(self.sums_all, self.sums_local)
        self._enq(PH_DECIDE)

    def linear_solve(self, lam):
        """(test hook) REDUCE + BACKSUB on the accepted state with the given lambda -> dx (M,75), dhalo (2,75)."""
        self.ctl[CTL_LAM] = float(lam)
        self._enq(PH_REDUCE)
        if self.world > 1:
            self._gather(self.gathered, self.payload)
        self._enq(PH_BACKSUB)
        return self.dx, self.dhalo

    def _exchange_halo_init(self, s):
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

    def exchange_us(self, reps=50):
        """Device time of the two per-attempt exchanges (CUDA events around all_gather x 2), microseconds."""
        torch = self.torch
        if self.world == 1:
            return 0.0
        for _ in range(5):
            self._gather(self.gathered, self.payload)
            self._gather(self.sums_all, self.sums_local)
        torch.cuda.synchronize(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            self._gather(self.gathered, self.payload)
            self._gather(self.sums_all, self.sums_local)
        e1.record()
        torch.cuda.synchronize(self.dev)
        return 1e3 * e0.elapsed_time(e1) / reps

    # -- the loop -------------------------------------------------------------------------------
    def solve(self, x0, max_iter=CHEETAH.lm_max_iter, lam0=CHEETAH.lm_lam0, tol_step=CHEETAH.lm_tol_step,
              tol_rel=CHEETAH.lm_tol_rel, max_attempts=CHEETAH.lm_max_attempts, verbose=False,
              tol_noise=CHEETAH.lm_tol_noise):
        """Stops when an accepted step moves no variable by more than tol_step, or lowers F by less than tol_rel * F,
        or when two consecutive REJECTED trial points differ from F by less than tol_noise * F (the measurement term
        is evaluated in fp32: ~6e-8 relative resolution per frame), or after max_attempts rejections in a row, or
        after max_iter accepted iterations.  (The reference stops IPOPT at tol = 1e-1, all_optimizations.py:511.)"""
        torch = self.torch
        N = self.N
        s = self.st[0]
        self.h.check_scene(self._scene_version, "FTESolver.solve")
        x0 = np.clip(np.asarray(x0, dtype=np.float64), self.lo.cpu().numpy(), self.hi.cpu().numpy())
        s["x_ext"].zero_()
        s["x_ext"][3:3 + N] = torch.as_tensor(x0).to(self.dev)
        s["x32"].copy_(s["x_ext"][3:3 + N].to(torch.float32))
        self._exchange_halo_init(s)
        ctl0 = np.zeros(CTL_SIZE)
        ctl0[CTL_LAM], ctl0[CTL_MAX_ITER], ctl0[CTL_MAX_ATTEMPTS] = lam0, max_iter, max_attempts
        ctl0[CTL_TOL_STEP], ctl0[CTL_TOL_REL], ctl0[CTL_HIST_CAP] = tol_step, tol_rel, self.HIST_CAP
        ctl0[CTL_TOL_NOISE], ctl0[CTL_DONE_AT] = tol_noise, -1.0
        if max_iter <= 0:
            ctl0[CTL_DONE], ctl0[CTL_STATUS] = 1, 3
        self.ctl.copy_(torch.from_numpy(ctl0))
        self.info.zero_()
        self._enq(PH_INIT_EVAL)
        if self.world > 1:
            self._gather(self.sums_all, self.sums_local)
        self._enq(PH_INIT_FINISH)
        torch.cuda.synchronize(self.dev)
        F0 = float(self.ctl_host[CTL_F])
        launches0 = self.h.launch_count
        cap = max_iter * max_attempts + 2
        evs = []                                      # one timing event after every enqueued attempt

        def enqueue(k):
            if self._graph is not None and k > 0:
                self._graph.replay()
                return
            self._attempt()                           # eager: the first attempt of every solve (= warm-up of the capture)
            if self.use_graph and self._graph is None and k == 0:
                torch.cuda.synchronize(self.dev)
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._attempt()
                    self._graph = g
                except Exception as e:                # capture is an optimisation, never a different code path
                    self._graph = None
                    self.use_graph = False
                    self._graph_error = repr(e)
                    torch.cuda.synchronize(self.dev)

        # The host stays ONE attempt ahead of the GPU: it enqueues attempt k, then waits for attempt k-1 only and looks
        # at the pinned mirror of the control block.  `done` is raised by the device at the same attempt index j on every
        # rank; every rank then makes sure it has enqueued exactly j + 2 attempts (an attempt contains collectives, so the
        # counts must match whatever each host happened to observe) - the surplus attempts are idle on the device.
        k = 0
        while k < cap and max_iter > 0:
            enqueue(k)
            evs.append(torch.cuda.Event(enable_timing=True))
            evs[k].record()
            k += 1
            if k >= 2:
                evs[k - 2].synchronize()              # attempt k-2 has finished
                done_at = int(self.ctl_host[CTL_DONE_AT])
                if done_at >= 0:
                    while k < done_at + 2:
                        enqueue(k)
                        k += 1
                    break
        torch.cuda.synchronize(self.dev)
        ctl = self.ctl.cpu().numpy()
        n_att = int(ctl[CTL_N_ATTEMPT])
        # device time of the graph-replayed attempts that did work (attempt 0 is eager + capture, attempts >= n_att idle)
        att_ms = [evs[i - 1].elapsed_time(evs[i]) for i in range(2, min(n_att, len(evs)))]
        hist = self.hist[:min(n_att, self.HIST_CAP)].cpu().numpy()
        if verbose and self.rank == 0:
            it = 0
            for r in hist:
                print(f"it {it:3d} lam {r[2]:9.3e} F {r[0]:16.6f} Ft {r[1]:16.6f} pred {r[6]:10.3e} rho {r[3]:7.3f} "
                      f"|dx|inf {r[4]:.2e}{' *' if r[5] else ''}")
                it += int(r[5])
        x = s["x_ext"][3:3 + N].cpu().numpy()
        history = [F0] + [float(r[1]) for r in hist if r[5]]
        status = int(ctl[CTL_STATUS])
        info = dict(F=float(ctl[CTL_F]), iters=int(ctl[CTL_ITERS]), n_eval=n_att + 1, n_solve=n_att, history=history,
                    lam=float(ctl[CTL_LAM]), converged=status == 1, status=STATUS.get(status, str(status)),
                    bcr_info=int(self.info.item()), launches=self.h.launch_count - self.n_launch0,
                    graph=self._graph is not None, attempts_enqueued=k,
                    collectives_per_attempt=0 if self.world == 1 else 2, host_syncs_per_attempt=0,
                    step=float(ctl[CTL_STEP]),
                    attempt_ms_steady=float(np.median(att_ms)) if att_ms else None)
        return x, info
