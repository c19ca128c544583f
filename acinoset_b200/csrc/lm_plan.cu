// FTE solve driver: one Levenberg-Marquardt attempt as stream-ordered phases, no host in the loop.
//
// Replaces `opt.solve(m)` (/root/reference/src/all_optimizations.py:503-524).  The outer loop state (lambda, objective,
// accept / reject, convergence) lives in a device control block; the phases below only enqueue kernels on the caller's
// stream, so a whole attempt can be captured into a CUDA graph (acinoset_b200/lm.py) and replayed until the pinned mirror
// of the control block says `done`.  Accept = copy trial -> accepted on the device (`lm_commit`), so every pointer in the
// graph is fixed.
#include <vector>

#include "handle.cuh"
#include "lm_common.cuh"

namespace acino {
cudaError_t launch_fte_eval(const SceneF& scene, int n_frames, const float* x, const float* meas, const float* w,
                            float* cost, float* g, float* H, cudaStream_t stream, int* sched);
cudaError_t launch_lm_prepare(int n_frames, long long frame0, long long ng, const double* x_ext, const float* g,
                              const double* sw, const double* lo, const double* hi, double* gtot, unsigned char* fixed,
                              double* cost_s, cudaStream_t s);
cudaError_t launch_lm_reduce(int n, const float* a0, const double* a1, const double* a2, const double* a3,
                             const double* m, double* out, double* ws, cudaStream_t s);
size_t lm_reduce_ws_bytes();
cudaError_t launch_bcr_factor(int n_elim, const int* elim, const double* D, const double* Lc, double* P, double* Q,
                              double* R, double* rhs, int* info, cudaStream_t s);
cudaError_t launch_bcr_update(int n_surv, const int* surv, double* D, double* Lc, const double* P, const double* Q,
                              double* rhs, cudaStream_t s);
cudaError_t launch_bcr_backsub(int n_elim, const int* elim, const double* R, const double* P, const double* Q,
                               const double* rhs, double* x, cudaStream_t s);
cudaError_t launch_l0_invert(const LmShard& sh, int n_elim, const int* elim, const float* H, const double* gtot,
                             const unsigned char* fixed, const double* sw, const double* ctl, double* W, double* rhs,
                             int* info, cudaStream_t s);
cudaError_t launch_l0_update(const LmShard& sh, int n_surv, const int* surv, const float* H, const double* gtot,
                             const unsigned char* fixed, const double* sw, const double* ctl, const double* W, double* D,
                             double* Lc, double* rhs, cudaStream_t s);
cudaError_t launch_l0_backsub(const LmShard& sh, int n_elim, const int* elim, const unsigned char* fixed, const double* sw,
                              const double* W, const double* rhs, double* x, cudaStream_t s);

constexpr size_t SB2 = (size_t)SBN * SBN;
static_assert(ACINO_LM_PAYLOAD == 4 * SBN * SBN + 4 * SBN, "payload layout");
static_assert(ACINO_LM_CTL == CTL_SIZE && ACINO_LM_SUMS == LM_SUMS && ACINO_LM_HIST == LM_HIST, "header / kernel constants");

// ---- interface exchange plumbing (world > 1) --------------------------------------------------------------------
// payload = [D_0 | D_{M-1} | Lc_0 | Lc_{M-1} | rhs_0 | rhs_{M-1} | frozen_0 | frozen_{M-1}] of the locally reduced chain
__global__ void lm_iface_pack_kernel(const int n_frames, const int M, const double* __restrict__ D,
                                     const double* __restrict__ Lc, const double* __restrict__ rhs,
                                     const unsigned char* __restrict__ fixed, double* __restrict__ payload) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ACINO_LM_PAYLOAD) return;
    const size_t last = (size_t)(M - 1);
    double v;
    if (i < 4 * (int)SB2) {
        const int which = i / (int)SB2, k = i - which * (int)SB2;
        v = which == 0 ? D[k] : which == 1 ? D[last * SB2 + k] : which == 2 ? Lc[k] : Lc[last * SB2 + k];
    } else {
        const int j = i - 4 * (int)SB2, which = j / SBN, k = j - which * SBN;
        if (which == 0) v = rhs[k];
        else if (which == 1) v = rhs[last * SBN + k];
        else {
            const long long n = (which == 2 ? 0 : 3 * (long long)last) + k / NA;
            v = (n < n_frames && fixed[(size_t)n * NA + k % NA]) ? 1.0 : 0.0;
        }
    }
    payload[i] = v;
}

// 2G-block interface chain from the gathered payloads: block 2r = rank r's first block, 2r + 1 = its last block.
// The coupling of rank r's first block to rank r-1's last block loses the columns rank r-1 froze.
__global__ void lm_iface_build_kernel(const int world, const double* __restrict__ gathered, double* __restrict__ cD,
                                      double* __restrict__ cLc, double* __restrict__ crhs) {
    const int blk = blockIdx.x;                 // chain block
    const int r = blk >> 1, which = blk & 1;
    const double* pl = gathered + (size_t)r * ACINO_LM_PAYLOAD;
    const double* prev_fixed = r > 0 ? gathered + (size_t)(r - 1) * ACINO_LM_PAYLOAD + 4 * SB2 + 3 * SBN : nullptr;
    for (int t = threadIdx.x; t < (int)SB2; t += blockDim.x) {
        cD[(size_t)blk * SB2 + t] = pl[(size_t)which * SB2 + t];
        double l = pl[(size_t)(2 + which) * SB2 + t];
        if (which == 0) l = r == 0 ? 0.0 : l * (1.0 - prev_fixed[t % SBN]);
        cLc[(size_t)blk * SB2 + t] = l;
    }
    for (int t = threadIdx.x; t < SBN; t += blockDim.x) crhs[(size_t)blk * SBN + t] = pl[4 * SB2 + (size_t)which * SBN + t];
}

// interface solution -> this rank's end blocks of dx and the halo steps (the neighbours' adjacent blocks)
__global__ void lm_iface_scatter_kernel(const int rank, const int world, const int M, const double* __restrict__ cx,
                                        double* __restrict__ dx, double* __restrict__ dhalo) {
    const int t = threadIdx.x;
    if (t >= SBN) return;
    dx[t] = cx[(size_t)(2 * rank) * SBN + t];
    dx[(size_t)(M - 1) * SBN + t] = cx[(size_t)(2 * rank + 1) * SBN + t];
    dhalo[t] = rank > 0 ? cx[(size_t)(2 * rank - 1) * SBN + t] : 0.0;
    dhalo[SBN + t] = rank < world - 1 ? cx[(size_t)(2 * rank + 2) * SBN + t] : 0.0;
}

// ---- trial point + model reduction ---------------------------------------------------------------------------------
// One warp per frame n in [-3, N + 3): lanes 0..24 own a parameter.  Interior frames: trial = clip(x + d), per-frame
// pred[n] = -g.d - 1/2 d^T H d - 1/2 sw (D3 d)^2 and step[n] = max |d| of the CLIPPED step; halo frames: the neighbour
// applies the same clamp to the same numbers, so only the trial state is written.
__device__ __forceinline__ double step_at(const int n, const int p, const int N, const double* __restrict__ dx,
                                          const double* __restrict__ dhalo) {
    if (n < 0) return dhalo[(3 + n) * NA + p];
    if (n >= N) return n < N + 3 ? dhalo[SBN + (n - N) * NA + p] : 0.0;
    return dx[(size_t)n * NA + p];
}

__global__ void lm_step2_kernel(const LmShard sh, const double* __restrict__ x_ext, const double* __restrict__ dx,
                                const double* __restrict__ dhalo, const double* __restrict__ gtot,
                                const float* __restrict__ H, const double* __restrict__ sw, const double* __restrict__ lo,
                                const double* __restrict__ hi, double* __restrict__ xt_ext, float* __restrict__ xt32,
                                double* __restrict__ pred, double* __restrict__ step) {
    const int N = sh.n_frames;
    const int warps_per_block = blockDim.x >> 5;
    const int n = blockIdx.x * warps_per_block + (threadIdx.x >> 5) - 3;
    const int p = threadIdx.x & 31;
    if (n >= N + 3) return;
    const bool act = p < NA;
    const size_t row = (size_t)(n + 3) * NA;
    double d = 0.0, xv = 0.0;
    if (act) {
        xv = x_ext[row + p];
        const double xt = fmin(fmax(xv + step_at(n, p, N, dx, dhalo), lo[p]), hi[p]);
        d = xt - xv;
        xt_ext[row + p] = xt;
        if (n >= 0 && n < N) xt32[(size_t)n * NA + p] = (float)xt;
    }
    if (n < 0 || n >= N) return;      // whole warp
    double hd = 0.0;
    for (int q = 0; q < NA; ++q) {
        const double dq = __shfl_sync(0xffffffffu, d, q);
        if (act) {
            const int lo_ = p < q ? p : q, hi_ = p < q ? q : p;
            hd = fma((double)H[(size_t)n * NU + upper_index(lo_, hi_)], dq, hd);
        }
    }
    double pr = 0.0, st = 0.0;
    if (act) {
        pr = -gtot[(size_t)n * NA + p] * d - 0.5 * d * hd;
        st = fabs(d);
        if (sh.frame0 + n >= 3) {
            double dd[4];
            dd[0] = d;
#pragma unroll
            for (int k = 1; k <= 3; ++k) {
                const double xk = x_ext[row - (size_t)k * NA + p];
                const double tk = fmin(fmax(xk + step_at(n - k, p, N, dx, dhalo), lo[p]), hi[p]);
                dd[k] = tk - xk;
            }
            const double d3 = dd[0] - 3.0 * dd[1] + 3.0 * dd[2] - dd[3];
            pr -= 0.5 * sw[p] * d3 * d3;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pr += __shfl_xor_sync(0xffffffffu, pr, o);
        st = fmax(st, __shfl_xor_sync(0xffffffffu, st, o));
    }
    if (p == 0) {
        pred[n] = pr;
        step[n] = st;
    }
}

// ---- control block -------------------------------------------------------------------------------------------------
// sums_all [world][LM_SUMS]: every rank adds the per-rank partials in rank order => identical bits everywhere.
__global__ void lm_init_finish_kernel(const int world, const double* __restrict__ sums_all, double* __restrict__ ctl) {
    if (threadIdx.x != 0) return;
    double F = 0.0;
    for (int r = 0; r < world; ++r) F += sums_all[r * LM_SUMS + 0] + sums_all[r * LM_SUMS + 1];
    ctl[CTL_F] = F;
}

__global__ void lm_decide_kernel(const int world, const double* __restrict__ sums_all, double* __restrict__ ctl,
                                 double* __restrict__ hist) {
    if (threadIdx.x != 0) return;
    ctl[CTL_ACCEPT] = 0.0;
    // CTL_N_ENQ counts every execution of this kernel (also the idle ones after `done`); CTL_DONE_AT is the execution
    // index at which `done` was raised.  Every rank raises it at the same index (identical inputs, identical bits), and
    // the hosts use it to enqueue the SAME number of attempts (= done_at + 2) however far ahead each of them is running -
    // an attempt contains collectives.
    const double enq = ctl[CTL_N_ENQ];
    ctl[CTL_N_ENQ] = enq + 1.0;
    if (ctl[CTL_DONE] != 0.0) return;
    double cost = 0.0, pred = 0.0, step = 0.0;
    for (int r = 0; r < world; ++r) {
        const double* s = sums_all + r * LM_SUMS;
        cost += s[0] + s[1];
        pred += s[2];
        step = fmax(step, s[4]);
    }
    const double F = ctl[CTL_F], Ft = cost, lam = ctl[CTL_LAM];
    const double rho = pred > 0.0 ? (F - Ft) / pred : -1.0;
    const int k = (int)ctl[CTL_N_ATTEMPT];
    ctl[CTL_N_ATTEMPT] = k + 1;
    ctl[CTL_FT] = Ft; ctl[CTL_PRED] = pred; ctl[CTL_STEP] = step; ctl[CTL_RHO] = rho;
    const bool accept = (Ft < F) && (rho > 1e-4);
    if (k < (int)ctl[CTL_HIST_CAP]) {
        double* h = hist + (size_t)k * LM_HIST;
        h[0] = F; h[1] = Ft; h[2] = lam; h[3] = rho; h[4] = step; h[5] = accept ? 1.0 : 0.0; h[6] = pred; h[7] = 0.0;
    }
    if (accept) {
        const double rel = (F - Ft) / fmax(fabs(F), 1e-30);
        ctl[CTL_REL] = rel;
        ctl[CTL_F] = Ft;
        ctl[CTL_ACCEPT] = 1.0;
        ctl[CTL_FAIL_STREAK] = 0.0;
        ctl[CTL_NOISE_STREAK] = 0.0;
        ctl[CTL_N_ACCEPT] += 1.0;
        ctl[CTL_ITERS] += 1.0;
        if (rho > 0.75) ctl[CTL_LAM] = fmax(lam / 3.0, 1e-12);
        else if (rho < 0.25) ctl[CTL_LAM] = lam * 2.0;
        if (step < ctl[CTL_TOL_STEP] || rel < ctl[CTL_TOL_REL]) {
            ctl[CTL_DONE] = 1.0;
            ctl[CTL_STATUS] = 1.0;                 // converged
        } else if (ctl[CTL_ITERS] >= ctl[CTL_MAX_ITER]) {
            ctl[CTL_DONE] = 1.0;
            ctl[CTL_STATUS] = 3.0;                 // iteration limit
        }
    } else {
        ctl[CTL_LAM] = lam * 4.0;
        const double fs = ctl[CTL_FAIL_STREAK] + 1.0;
        ctl[CTL_FAIL_STREAK] = fs;
        // The objective is evaluated in fp32: a rejected step whose objective differs from the accepted one by less
        // than the evaluation's resolution carries no information.  Two of those in a row = converged.
        const double ns = (fabs(Ft - F) <= ctl[CTL_TOL_NOISE] * fabs(F)) ? ctl[CTL_NOISE_STREAK] + 1.0 : 0.0;
        ctl[CTL_NOISE_STREAK] = ns;
        if (ns >= 2.0) {
            ctl[CTL_DONE] = 1.0;
            ctl[CTL_STATUS] = 1.0;                 // converged (objective change below the evaluation resolution)
        } else if (fs >= ctl[CTL_MAX_ATTEMPTS]) {
            ctl[CTL_ITERS] += 1.0;
            ctl[CTL_DONE] = 1.0;
            ctl[CTL_STATUS] = 2.0;                 // no acceptable step
        }
    }
    if (ctl[CTL_DONE] != 0.0) ctl[CTL_DONE_AT] = enq;
}

// accepted <- trial when the last attempt was accepted (16-byte vectors; every buffer is a separate allocation)
struct CommitJob {
    void* dst;
    const void* src;
    size_t bytes;
};
struct CommitJobs {
    CommitJob j[4];
};
__global__ void lm_commit_kernel(const double* __restrict__ ctl, const CommitJobs jobs) {
    if (ctl[CTL_ACCEPT] == 0.0) return;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t n16 = jobs.j[k].bytes >> 4;
        uint4* d = reinterpret_cast<uint4*>(jobs.j[k].dst);
        const uint4* s = reinterpret_cast<const uint4*>(jobs.j[k].src);
        for (size_t i = tid; i < n16; i += nt) d[i] = s[i];
        const size_t tail = jobs.j[k].bytes & 15;
        if (tid < tail) {
            reinterpret_cast<unsigned char*>(jobs.j[k].dst)[(n16 << 4) + tid] =
                reinterpret_cast<const unsigned char*>(jobs.j[k].src)[(n16 << 4) + tid];
        }
    }
}

}  // namespace acino

struct acino_lm_plan {
    acino_lm_desc d;
    int* tile_sched = nullptr;          // two {ticket, finished} pairs for the plan's two fte_eval launches (captured in its graph)
    std::vector<int> lvl, clvl;         // (n_elim, n_surv) per level: local chain / interface chain
};

extern "C" {

int acino_lm_desc_size(void) { return (int)sizeof(acino_lm_desc); }

int acino_lm_plan_create(acino_handle* h, const acino_lm_desc* desc, acino_lm_plan** out) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_lm_plan_create: NULL handle");
    if (!desc || !out) return fail(h, ACINO_ERR_ARG, "acino_lm_plan_create: NULL argument");
    const acino_lm_desc& d = *desc;
    if (d.n_frames < 1 || d.n_blocks * 3 < d.n_frames || d.world < 1 || d.rank < 0 || d.rank >= d.world)
        return fail(h, ACINO_ERR_ARG, "acino_lm_plan_create: bad shard description");
    if (d.world > 1 && d.n_blocks < 2) return fail(h, ACINO_ERR_ARG, "acino_lm_plan_create: a rank needs at least 2 super-blocks");
    if (!d.meas || !d.w || !d.sw || !d.lo || !d.hi || !d.pred || !d.step || !d.D || !d.Lc || !d.P || !d.Q || !d.R || !d.rhs || !d.dx ||
        !d.dhalo || !d.info || !d.sums_local || !d.sums_all || !d.ctl || !d.ctl_host || !d.hist)
        return fail(h, ACINO_ERR_ARG, "acino_lm_plan_create: NULL buffer");
    for (int i = 0; i < 2; ++i)
        if (!d.x_ext[i] || !d.x32[i] || !d.cost[i] || !d.g[i] || !d.H[i] || !d.gtot[i] || !d.fixed[i] || !d.cost_s[i])
            return fail(h, ACINO_ERR_ARG, "acino_lm_plan_create: NULL state buffer");
    if ((d.n_elim0 > 0 && !d.elim0) || (d.n_surv0 > 0 && !d.surv0) || (d.n_levels > 0 && (!d.level_counts || !d.sched)))
        return fail(h, ACINO_ERR_ARG, "acino_lm_plan_create: NULL schedule");
    if (d.world > 1 && (!d.payload || !d.gathered || !d.cD || !d.cLc || !d.cP || !d.cQ || !d.cR || !d.crhs || !d.cx ||
                        d.n_clevels < 1 || !d.clevel_counts || !d.csched))
        return fail(h, ACINO_ERR_ARG, "acino_lm_plan_create: NULL interface-chain buffer");
    if (!h->have_cams) return fail(h, ACINO_ERR_STATE, "acino_lm_plan_create: cameras not set");
    CK(cudaSetDevice(h->device));
    if (!h->red_ws) {
        CK(cudaMalloc((void**)&h->red_ws, lm_reduce_ws_bytes()));
        CK(cudaMemset(h->red_ws, 0, lm_reduce_ws_bytes()));
        CK(cudaDeviceSynchronize());
    }
    acino_lm_plan* p = new acino_lm_plan();
    p->d = d;
    if (cudaMalloc((void**)&p->tile_sched, 4 * sizeof(int)) != cudaSuccess || cudaMemset(p->tile_sched, 0, 4 * sizeof(int)) != cudaSuccess) {
        delete p;
        return fail(h, ACINO_ERR_CUDA, "acino_lm_plan_create: cannot allocate the tile-schedule counters");
    }
    p->lvl.assign(d.level_counts, d.level_counts + 2 * (size_t)d.n_levels);
    if (d.world > 1) p->clvl.assign(d.clevel_counts, d.clevel_counts + 2 * (size_t)d.n_clevels);
    p->d.level_counts = nullptr;
    p->d.clevel_counts = nullptr;
    *out = p;
    return ACINO_OK;
}

int acino_lm_plan_destroy(acino_lm_plan* plan) {
    if (plan && plan->tile_sched) cudaFree(plan->tile_sched);
    delete plan;
    return ACINO_OK;
}

// forward elimination / back-substitution of the dense levels of a chain
static int chain_reduce(acino_handle* h, const std::vector<int>& lvl, const int32_t* sched, double* D, double* Lc, double* P,
                        double* Q, double* R, double* rhs, int32_t* info, cudaStream_t s) {
    const int32_t* ptr = sched;
    for (size_t l = 0; l < lvl.size() / 2; ++l) {
        const int ne = lvl[2 * l], ns = lvl[2 * l + 1];
        CK(launch_bcr_factor(ne, ptr, D, Lc, P, Q, R, rhs, info, s));
        CK(launch_bcr_update(ns, ptr + 3 * (size_t)ne, D, Lc, P, Q, rhs, s));
        h->launches += (ne > 0) + (ns > 0);
        ptr += 3 * (size_t)(ne + ns);
    }
    return ACINO_OK;
}
static int chain_backsub(acino_handle* h, const std::vector<int>& lvl, const int32_t* sched, const double* R, const double* P,
                         const double* Q, const double* rhs, double* x, cudaStream_t s) {
    std::vector<const int32_t*> ptrs(lvl.size() / 2);
    const int32_t* ptr = sched;
    for (size_t l = 0; l < lvl.size() / 2; ++l) {
        ptrs[l] = ptr;
        ptr += 3 * (size_t)(lvl[2 * l] + lvl[2 * l + 1]);
    }
    for (size_t l = lvl.size() / 2; l-- > 0;) {
        CK(launch_bcr_backsub(lvl[2 * l], ptrs[l], R, P, Q, rhs, x, s));
        h->launches += lvl[2 * l] > 0;
    }
    return ACINO_OK;
}

static int eval_state(acino_handle* h, const acino_lm_desc& d, int i, bool with_step, cudaStream_t s, int* tile_sched) {
    CK(launch_fte_eval(h->scene, d.n_frames, d.x32[i], d.meas, d.w, d.cost[i], d.g[i], d.H[i], s, tile_sched + 2 * i));
    CK(launch_lm_prepare(d.n_frames, d.frame0, d.n_global, d.x_ext[i], d.g[i], d.sw, d.lo, d.hi, d.gtot[i], d.fixed[i],
                         d.cost_s[i], s));
    CK(launch_lm_reduce(d.n_frames, d.cost[i], d.cost_s[i], with_step ? d.pred : nullptr, nullptr,
                        with_step ? d.step : nullptr, d.sums_local, h->red_ws, s));
    h->launches += 3;
    return ACINO_OK;
}

int acino_lm_enqueue(acino_handle* h, acino_lm_plan* plan, int phase, void* cuda_stream) {
    if (!h) return fail(nullptr, ACINO_ERR_ARG, "acino_lm_enqueue: NULL handle");
    if (!plan) return fail(h, ACINO_ERR_ARG, "acino_lm_enqueue: NULL plan");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const acino_lm_desc& d = plan->d;
    const LmShard sh{d.n_frames, (long long)d.frame0, (long long)d.n_global};
    const int N = d.n_frames, M = d.n_blocks;
    int rc;
    switch (phase) {
    case ACINO_LM_INIT_EVAL:
        return eval_state(h, d, 0, false, s, plan->tile_sched);
    case ACINO_LM_INIT_FINISH:
        lm_init_finish_kernel<<<1, 32, 0, s>>>(d.world, d.sums_all, d.ctl);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(d.ctl_host, d.ctl, CTL_SIZE * sizeof(double), cudaMemcpyDeviceToHost, s));
        h->launches += 1;
        return ACINO_OK;
    case ACINO_LM_REDUCE:
        CK(launch_l0_invert(sh, d.n_elim0, d.elim0, d.H[0], d.gtot[0], d.fixed[0], d.sw, d.ctl, d.P, d.rhs, d.info, s));
        CK(launch_l0_update(sh, d.n_surv0, d.surv0, d.H[0], d.gtot[0], d.fixed[0], d.sw, d.ctl, d.P, d.D, d.Lc, d.rhs, s));
        h->launches += (d.n_elim0 > 0) + (d.n_surv0 > 0);
        rc = chain_reduce(h, plan->lvl, d.sched, d.D, d.Lc, d.P, d.Q, d.R, d.rhs, d.info, s);
        if (rc) return rc;
        if (d.world > 1) {
            lm_iface_pack_kernel<<<(ACINO_LM_PAYLOAD + 255) / 256, 256, 0, s>>>(N, M, d.D, d.Lc, d.rhs, d.fixed[0], d.payload);
            CK(cudaGetLastError());
            h->launches += 1;
        }
        return ACINO_OK;
    case ACINO_LM_BACKSUB:
        if (d.world > 1) {
            lm_iface_build_kernel<<<2 * d.world, 256, 0, s>>>(d.world, d.gathered, d.cD, d.cLc, d.crhs);
            CK(cudaGetLastError());
            h->launches += 1;
            rc = chain_reduce(h, plan->clvl, d.csched, d.cD, d.cLc, d.cP, d.cQ, d.cR, d.crhs, d.info, s);
            if (rc) return rc;
            rc = chain_backsub(h, plan->clvl, d.csched, d.cR, d.cP, d.cQ, d.crhs, d.cx, s);
            if (rc) return rc;
            lm_iface_scatter_kernel<<<1, 96, 0, s>>>(d.rank, d.world, M, d.cx, d.dx, d.dhalo);
            CK(cudaGetLastError());
            h->launches += 1;
        }
        rc = chain_backsub(h, plan->lvl, d.sched, d.R, d.P, d.Q, d.rhs, d.dx, s);
        if (rc) return rc;
        CK(launch_l0_backsub(sh, d.n_elim0, d.elim0, d.fixed[0], d.sw, d.P, d.rhs, d.dx, s));
        h->launches += d.n_elim0 > 0;
        return ACINO_OK;
    case ACINO_LM_TRIAL: {
        const int wpb = 8;
        lm_step2_kernel<<<(N + 6 + wpb - 1) / wpb, wpb * 32, 0, s>>>(sh, d.x_ext[0], d.dx, d.dhalo, d.gtot[0], d.H[0], d.sw, d.lo,
                                                                      d.hi, d.x_ext[1], d.x32[1], d.pred, d.step);
        CK(cudaGetLastError());
        h->launches += 1;
        return eval_state(h, d, 1, true, s, plan->tile_sched);
    }
    case ACINO_LM_DECIDE: {
        lm_decide_kernel<<<1, 32, 0, s>>>(d.world, d.sums_all, d.ctl, d.hist);
        CK(cudaGetLastError());
        CommitJobs jobs;
        jobs.j[0] = {d.x_ext[0], d.x_ext[1], (size_t)(N + 6) * NA * sizeof(double)};
        jobs.j[1] = {d.H[0], d.H[1], (size_t)N * NU * sizeof(float)};
        jobs.j[2] = {d.gtot[0], d.gtot[1], (size_t)N * NA * sizeof(double)};
        jobs.j[3] = {d.fixed[0], d.fixed[1], (size_t)N * NA};
        const size_t work = ((size_t)N * NU * sizeof(float)) >> 4;
        int grid = (int)((work + 255) / 256);
        grid = grid < 1 ? 1 : (grid > 148 * 8 ? 148 * 8 : grid);
        lm_commit_kernel<<<grid, 256, 0, s>>>(d.ctl, jobs);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(d.ctl_host, d.ctl, CTL_SIZE * sizeof(double), cudaMemcpyDeviceToHost, s));
        h->launches += 2;
        return ACINO_OK;
    }
    default:
        return fail(h, ACINO_ERR_ARG, "acino_lm_enqueue: unknown phase");
    }
}

}  // extern "C"
