// Levenberg-Marquardt glue kernels for the FTE solve (fp64 state, fp32 data blocks from fte_eval).
//
// Replaces the NLP the reference hands to IPOPT (/root/reference/src/all_optimizations.py):
//   backwards_euler_pos / _vel / constant_acc :369-391 + model term of obj :490-492
//       => smoothness cost  sum_{n>=3,p} q_p ((x_n - 3x_{n-1} + 3x_{n-2} - x_{n-3}) / Ts^2)^2
//   21 pose bounds :403-483  => box projection + frozen (active) variables
// Kernels: lm_prepare (smoothness gradient/cost, active set), lm_assemble (75x75 super-blocks of
// B + lam diag(B), couplings, right-hand side), lm_step (projected trial point, model reduction),
// lm_reduce (fixed-order fp64 sums).  Frame indices are GLOBAL (frame0 + local) so that a rank
// holding a contiguous shard with 3-frame halos builds exactly its rows of the global system.
#include "lm_common.cuh"

namespace acino {

// thread per (frame, parameter).  x_ext has 3 halo frames on each side: row (n + 3) is local frame n.
__global__ void lm_prepare_kernel(const int n_frames, const long long frame0, const long long ng,
                                  const double* __restrict__ x_ext, const float* __restrict__ g,
                                  const double* __restrict__ sw, const double* __restrict__ lo,
                                  const double* __restrict__ hi, double* __restrict__ gtot,
                                  unsigned char* __restrict__ fixed, double* __restrict__ cost_s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n_frames * NA;
    const int n = live ? i / NA : 0, p = live ? i - n * NA : 0;
    const long long gn = frame0 + n;
    const double* xc = x_ext + (size_t)(n + 3) * NA + p;
    double cs = 0.0;
    if (live) {
        // smoothness gradient: sw_p * sum_b (D3^T D3)[gn][b] x_b
        double gs = 0.0;
#pragma unroll
        for (int k = -3; k <= 3; ++k) {
            const double c = k >= 0 ? band_coef(gn, k, ng) : band_coef(gn + k, -k, ng);
            if (c != 0.0) gs = fma(c, xc[k * NA], gs);
        }
        gs *= sw[p];
        const double gt = (double)g[i] + gs;
        gtot[i] = gt;
        const double xv = xc[0];
        fixed[i] = ((xv <= lo[p] && gt > 0.0) || (xv >= hi[p] && gt < 0.0)) ? 1 : 0;
        // smoothness cost of the third difference ENDING at this frame (global m = gn >= 3)
        if (gn >= 3) {
            const double d3 = xc[0] - 3.0 * xc[-NA] + 3.0 * xc[-2 * NA] - xc[-3 * NA];
            cs = 0.5 * sw[p] * d3 * d3;      // sw = 2 q / Ts^4  ->  q (d3 / Ts^2)^2
        }
    }
    // reduce the 25 parameters of a frame: fixed order, by thread p == 0
    __shared__ double sc[256];
    sc[threadIdx.x] = cs;
    __syncthreads();
    if (live && p == 0) {      // blockDim = 250 = 10 whole frames: a frame never straddles CTAs
        double s = 0.0;
        for (int q = 0; q < NA; ++q) s += sc[threadIdx.x + q];
        cost_s[n] = s;
    }
}

// one CTA per super-block (3 frames).  D [M][75][75], Lc [M][75][75] (coupling to the previous
// super-block, global frames frame0 + 3i - 3 ..), rhs [M][75].
__global__ void __launch_bounds__(256)
lm_assemble_kernel(const int n_frames, const long long frame0, const long long ng, const float* __restrict__ H,
                   const double* __restrict__ gtot, const unsigned char* __restrict__ fixed,
                   const double* __restrict__ sw, const double lambda, double* __restrict__ D,
                   double* __restrict__ Lc, double* __restrict__ rhs) {
    const int sb = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ unsigned char sfix[2 * SBN];     // [0..74] this block, [75..149] previous block
    __shared__ double sdiag[SBN];               // diag(B) of this block
    for (int t = tid; t < 2 * SBN; t += 256) {
        const int which = t / SBN, r = t - which * SBN;
        const int n = 3 * (sb - which) + r / NA;
        sfix[t] = (n >= 0 && n < n_frames) ? fixed[(size_t)n * NA + r % NA] : 0;
    }
    for (int t = tid; t < SBN; t += 256) {
        const int a = t / NA, p = t - a * NA;
        const int n = 3 * sb + a;
        double d = 1.0;
        if (n < n_frames) d = (double)H[(size_t)n * NU + upper_index(p, p)] + band_coef(frame0 + n, 0, ng) * sw[p];
        sdiag[t] = d;
    }
    __syncthreads();
    double* Db = D + (size_t)sb * SBN * SBN;
    double* Lb = Lc + (size_t)sb * SBN * SBN;
    for (int t = tid; t < SBN * SBN; t += 256) {
        const int r = t / SBN, c = t - r * SBN;
        const int a = r / NA, p = r - a * NA, b = c / NA, q = c - b * NA;
        const int na = 3 * sb + a, nb = 3 * sb + b;
        // ---- diagonal super-block
        double v = 0.0;
        const bool va = na < n_frames, vb = nb < n_frames;
        if (!va || !vb || sfix[r] || sfix[c]) {
            v = (r == c) ? 1.0 : 0.0;                       // padding frame or frozen variable
        } else if (a == b) {
            const int lo_ = p < q ? p : q, hi_ = p < q ? q : p;
            v = (double)H[(size_t)na * NU + upper_index(lo_, hi_)];
            if (p == q) v = sdiag[r] * (1.0 + lambda);       // Marquardt: B_pp + lam B_pp
        } else if (p == q) {
            const int k = a > b ? a - b : b - a;
            v = band_coef(frame0 + (a < b ? na : nb), k, ng) * sw[p];
        }
        Db[t] = v;
        // ---- coupling to the previous super-block: rows = this block, cols = frames 3(sb-1)+b
        double u = 0.0;
        const int nbp = 3 * (sb - 1) + b;
        const int kk = na - nbp;                             // 1..5
        const bool prev_valid = (frame0 + nbp) >= 0 && (sb > 0 || frame0 > 0);
        if (p == q && kk <= 3 && va && prev_valid && !sfix[r] && !sfix[SBN + c])
            u = band_coef(frame0 + nbp, kk, ng) * sw[p];
        Lb[t] = u;
    }
    for (int t = tid; t < SBN; t += 256) {
        const int n = 3 * sb + t / NA;
        rhs[(size_t)sb * SBN + t] = (n < n_frames && !sfix[t]) ? -gtot[(size_t)n * NA + t % NA] : 0.0;
    }
}

// trial point + model reduction.  thread per (frame, parameter); d_ext [(N+6)][25] is the solver's
// step with halos (0 beyond the global ends).  Writes x_trial (fp64, into x_trial_ext interior),
// x_trial32, and per-frame partials: pred[n] = -g.d - 1/2 d^T H d - 1/2 sw (D3 d)^2, step[n] = max |d|.
__global__ void lm_step_kernel(const int n_frames, const long long frame0, const long long ng,
                               const double* __restrict__ x_ext, const double* __restrict__ d_ext,
                               const double* __restrict__ gtot, const float* __restrict__ H,
                               const double* __restrict__ sw, const double* __restrict__ lo,
                               const double* __restrict__ hi, double* __restrict__ xt_ext,
                               float* __restrict__ xt32, double* __restrict__ pred, double* __restrict__ step) {
    // one warp per frame: lanes 0..24 own a parameter
    const int warps_per_block = blockDim.x >> 5;
    const int n = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int p = threadIdx.x & 31;
    if (n >= n_frames) return;
    const bool act = p < NA;
    const size_t row = (size_t)(n + 3) * NA;
    double d = 0.0, xv = 0.0;
    if (act) {
        xv = x_ext[row + p];
        const double xt = fmin(fmax(xv + d_ext[row + p], lo[p]), hi[p]);
        d = xt - xv;
        xt_ext[row + p] = xt;
        xt32[(size_t)n * NA + p] = (float)xt;
    }
    // H d for this frame: (H d)_p = sum_q H[p][q] d_q
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
        const long long gn = frame0 + n;
        if (gn >= 3) {
            // third difference of the CLIPPED step ending at this frame; neighbours' clipped steps are
            // recomputed from x/d (cheap) so that no second pass is needed
            double dd[4];
            dd[0] = d;
#pragma unroll
            for (int k = 1; k <= 3; ++k) {
                const double xk = x_ext[row - (size_t)k * NA + p];
                const double tk = fmin(fmax(xk + d_ext[row - (size_t)k * NA + p], lo[p]), hi[p]);
                dd[k] = tk - xk;
            }
            const double d3 = dd[0] - 3.0 * dd[1] + 3.0 * dd[2] - dd[3];
            pr -= 0.5 * sw[p] * d3 * d3;       // S = sw D3^T D3  =>  1/2 d^T S d = 1/2 sw sum (D3 d)^2
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

// fixed-order fp64 reduction of up to 4 arrays (sum) + 1 array (max): single CTA, 1024 threads.
// out[0..3] = sums (a0 is float), out[4] = max of m.  Two stages, both in a FIXED order so the result is
// bit-reproducible: CTA b reduces elements b*1024 + t, (b + G)*1024 + t, ... into partial[b][5]; the last CTA to
// finish (ticket counter) adds the G partials in order b = 0..G-1.  ws: [G][5] doubles followed by the counter.
constexpr int RED_MAX_CTAS = 148;
__global__ void __launch_bounds__(1024)
lm_reduce_kernel(const int n, const float* __restrict__ a0, const double* __restrict__ a1, const double* __restrict__ a2,
                 const double* __restrict__ a3, const double* __restrict__ m, double* __restrict__ out,
                 double* __restrict__ ws) {
    __shared__ double s[5][1024];
    __shared__ bool last;
    const int G = gridDim.x;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0, vm = 0;
    for (size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x; i < (size_t)n; i += (size_t)G * 1024) {
        if (a0) v0 += (double)a0[i];
        if (a1) v1 += a1[i];
        if (a2) v2 += a2[i];
        if (a3) v3 += a3[i];
        if (m) vm = fmax(vm, m[i]);
    }
    s[0][threadIdx.x] = v0; s[1][threadIdx.x] = v1; s[2][threadIdx.x] = v2; s[3][threadIdx.x] = v3; s[4][threadIdx.x] = vm;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s[k][threadIdx.x] += s[k][threadIdx.x + o];
            s[4][threadIdx.x] = fmax(s[4][threadIdx.x], s[4][threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (G == 1) {
        if (threadIdx.x < 5) out[threadIdx.x] = s[threadIdx.x][0];
        return;
    }
    unsigned* counter = reinterpret_cast<unsigned*>(ws + RED_MAX_CTAS * 5);
    if (threadIdx.x < 5) ws[blockIdx.x * 5 + threadIdx.x] = s[threadIdx.x][0];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(counter, 1u);
        last = ticket == (unsigned)(G - 1);
        if (last) *counter = 0;            // re-armed for the next launch (stream-ordered)
    }
    __syncthreads();
    if (last && threadIdx.x < 5) {
        __threadfence();
        const volatile double* w = ws;
        double acc = 0.0;
        for (int b = 0; b < G; ++b) acc = threadIdx.x < 4 ? acc + w[b * 5 + threadIdx.x] : fmax(acc, w[b * 5 + 4]);
        out[threadIdx.x] = acc;
    }
}

cudaError_t launch_lm_prepare(int n_frames, long long frame0, long long ng, const double* x_ext, const float* g,
                              const double* sw, const double* lo, const double* hi, double* gtot, unsigned char* fixed,
                              double* cost_s, cudaStream_t s) {
    if (n_frames <= 0) return cudaSuccess;
    // 250 threads = 10 whole frames per CTA, so the per-frame cost reduction never straddles CTAs
    const int threads = 250, total = n_frames * NA;
    lm_prepare_kernel<<<(total + threads - 1) / threads, threads, 0, s>>>(n_frames, frame0, ng, x_ext, g, sw, lo, hi, gtot,
                                                                       fixed, cost_s);
    return cudaGetLastError();
}

cudaError_t launch_lm_assemble(int n_frames, long long frame0, long long ng, int n_blocks, const float* H,
                               const double* gtot, const unsigned char* fixed, const double* sw, double lambda, double* D,
                               double* Lc, double* rhs, cudaStream_t s) {
    if (n_blocks <= 0) return cudaSuccess;
    lm_assemble_kernel<<<n_blocks, 256, 0, s>>>(n_frames, frame0, ng, H, gtot, fixed, sw, lambda, D, Lc, rhs);
    return cudaGetLastError();
}

cudaError_t launch_lm_step(int n_frames, long long frame0, long long ng, const double* x_ext, const double* d_ext,
                           const double* gtot, const float* H, const double* sw, const double* lo, const double* hi,
                           double* xt_ext, float* xt32, double* pred, double* step, cudaStream_t s) {
    if (n_frames <= 0) return cudaSuccess;
    const int wpb = 8;
    lm_step_kernel<<<(n_frames + wpb - 1) / wpb, wpb * 32, 0, s>>>(n_frames, frame0, ng, x_ext, d_ext, gtot, H, sw, lo, hi,
                                                                   xt_ext, xt32, pred, step);
    return cudaGetLastError();
}

size_t lm_reduce_ws_bytes() { return (RED_MAX_CTAS * 5 + 1) * sizeof(double); }

cudaError_t launch_lm_reduce(int n, const float* a0, const double* a1, const double* a2, const double* a3,
                             const double* m, double* out, double* ws, cudaStream_t s) {
    int G = (n + 8191) / 8192;             // >= 8 elements per thread before a second CTA pays off
    G = G < 1 ? 1 : (G > RED_MAX_CTAS ? RED_MAX_CTAS : G);
    if (!ws) G = 1;
    lm_reduce_kernel<<<G, 1024, 0, s>>>(n, a0, a1, a2, a3, m, out, ws);
    return cudaGetLastError();
}

}  // namespace acino
