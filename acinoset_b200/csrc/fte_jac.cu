// fte_jac: dense measurement Jacobian  d (u,v)[c][l] / d x[25]  of every frame (the "dense-J-out" variant of
// the reprojection path).
//
// Replaces (reference, /root/reference/src/all_optimizations.py): the EKF's h_function (:615-621, FK +
// project_points_fisheye per camera) and numerical_jacobian (:634-649: 25 forward-difference perturbations
// of h per camera per frame, eps = 1e-3) used at :800-806.  Here the Jacobian is analytic:
//     d uv / d x = [d uv / d Xw] [I | omega_a x (p_l - pivot_a)]      (closed form)
// with the world-frame twists (omega_a, v_a = pivot_a x omega_a) of cheetah_fk.cuh, and
//     Ju . (omega x p + v) = omega . (p x Ju) + v . Ju               (one cross product per row, 6 FMA per entry)
//
// Layout: one CTA = 8 frames, thread <-> (frame, marker), loop over cameras.  Per camera the CTA's
// 8 x 20 x 50 Jacobian entries are staged in shared memory in their global layout (per frame 4000 contiguous
// bytes) and written out by bulk async stores (TMA, one per frame, issued by one thread); the stage is double
// buffered so the stores of camera c overlap the arithmetic of camera c+1.  The kernel is bound by the
// 24 KB/frame of Jacobian it writes to HBM.
#include "acino_common.cuh"
#include "cheetah_fk.cuh"

namespace acino {

constexpr int FTJ = 8;
constexpr int JROW = 2 * NA;        // 50 entries per (camera, marker): d u / d x[25], d v / d x[25]
constexpr int PER_F = NL * JROW;    // 1000 floats = 4000 bytes per (frame, camera)

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// bit a set <=> angle slot a moves marker l (its joint is an ancestor-or-self of the marker's joint)
struct MarkerMasks {
    unsigned m[NL];
};
constexpr MarkerMasks make_marker_masks() {
    MarkerMasks t{};
    for (int l = 0; l < NL; ++l)
        for (int a = 0; a < NANG; ++a)
            if (joint_is_anc(k_angle_joint[a], k_marker_joint[l])) t.m[l] |= 1u << a;
    return t;
}
__constant__ MarkerMasks c_marker_masks = make_marker_masks();

struct SmemJ {
    __align__(16) float stage[2][FTJ][PER_F];
    float x[FTJ][NA];
    float2 sc[FTJ][NANG];
    float p[FTJ][NL][3];
    __align__(16) float tau[FTJ][TAUF];
};

__global__ void __launch_bounds__(FTJ * NL)
fte_jac_kernel(const __grid_constant__ SceneF scene, const int n_frames, const int bulk_ok, const float* __restrict__ xg,
               float* __restrict__ uv_out, float* __restrict__ J_out) {
    constexpr int NT = FTJ * NL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemJ& S = *reinterpret_cast<SmemJ*>(smem_raw);
    const int tid = threadIdx.x;
    const int f0 = blockIdx.x * FTJ;
    const int nf = min(FTJ, n_frames - f0);
    const int C = scene.n_cams;
    for (int i = tid; i < FTJ * NA; i += NT) {
        const int f = i / NA;
        (&S.x[0][0])[i] = (f < nf) ? xg[(size_t)f0 * NA + i] : 0.f;
    }
    __syncthreads();
    for (int t = tid; t < FTJ * NANG; t += NT) {
        const int a = t / FTJ, f = t - a * FTJ;
        float sn, cs;
        sincosf(S.x[f][3 + a], &sn, &cs);
        S.sc[f][a] = make_float2(sn, cs);
    }
    __syncthreads();
    if (tid < FTJ) {
        FkWriter w{&S.p[tid][0][0], &S.tau[tid][0]};
        cheetah_fk(S.sc[tid], w);
    }
    __syncthreads();
    const int f = tid / NL;
    const int l = tid - f * NL;
    const float px = S.p[f][l][0], py = S.p[f][l][1], pz = S.p[f][l][2];     // relative to the head point
    const float wx = S.x[f][0] + px, wy = S.x[f][1] + py, wz = S.x[f][2] + pz;
    const unsigned mask = c_marker_masks.m[l];
    const float* tau = &S.tau[f][0];
    for (int c = 0; c < C; ++c) {
        const CamF& cam = scene.cam[c];
        const float xc = fmaf(cam.R[0], wx, fmaf(cam.R[1], wy, fmaf(cam.R[2], wz, cam.t[0])));
        const float yc = fmaf(cam.R[3], wx, fmaf(cam.R[4], wy, fmaf(cam.R[5], wz, cam.t[1])));
        const float zc = fmaf(cam.R[6], wx, fmaf(cam.R[7], wy, fmaf(cam.R[8], wz, cam.t[2])));
        ProjOut<float> pr;
        fisheye_cam<float, true>(xc, yc, zc, cam.fx, cam.fy, cam.D[0], cam.D[1], cam.D[2], cam.D[3], pr);
        if (uv_out && f < nf)
            reinterpret_cast<float2*>(uv_out)[((size_t)(f0 + f) * C + c) * NL + l] = make_float2(pr.u + cam.cx, pr.v + cam.cy);
        if (!J_out) continue;
        // the stage buffer of camera c-2 must have been read by its bulk stores before it is overwritten
        if (bulk_ok && c >= 2) {
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
        }
        float* row = &S.stage[c & 1][f][l * JROW];
        // world-frame rows  Ju = ju^T R,  Jv = jv^T R
        const float ju0 = fmaf(pr.ju[0], cam.R[0], fmaf(pr.ju[1], cam.R[3], pr.ju[2] * cam.R[6]));
        const float ju1 = fmaf(pr.ju[0], cam.R[1], fmaf(pr.ju[1], cam.R[4], pr.ju[2] * cam.R[7]));
        const float ju2 = fmaf(pr.ju[0], cam.R[2], fmaf(pr.ju[1], cam.R[5], pr.ju[2] * cam.R[8]));
        const float jv0 = fmaf(pr.jv[0], cam.R[0], fmaf(pr.jv[1], cam.R[3], pr.jv[2] * cam.R[6]));
        const float jv1 = fmaf(pr.jv[0], cam.R[1], fmaf(pr.jv[1], cam.R[4], pr.jv[2] * cam.R[7]));
        const float jv2 = fmaf(pr.jv[0], cam.R[2], fmaf(pr.jv[1], cam.R[5], pr.jv[2] * cam.R[8]));
        // m = p x J
        const float mu0 = py * ju2 - pz * ju1, mu1 = pz * ju0 - px * ju2, mu2 = px * ju1 - py * ju0;
        const float mv0 = py * jv2 - pz * jv1, mv1 = pz * jv0 - px * jv2, mv2 = px * jv1 - py * jv0;
        row[0] = ju0; row[1] = ju1; row[2] = ju2;
        row[NA + 0] = jv0; row[NA + 1] = jv1; row[NA + 2] = jv2;
#pragma unroll
        for (int a = 0; a < NANG; ++a) {
            const float4 t0 = *reinterpret_cast<const float4*>(tau + a * TAU_STRIDE);
            const float2 t1 = *reinterpret_cast<const float2*>(tau + a * TAU_STRIDE + 4);
            const bool on = (mask >> a) & 1u;
            const float du = fmaf(t0.x, mu0, fmaf(t0.y, mu1, fmaf(t0.z, mu2, fmaf(t0.w, ju0, fmaf(t1.x, ju1, t1.y * ju2)))));
            const float dv = fmaf(t0.x, mv0, fmaf(t0.y, mv1, fmaf(t0.z, mv2, fmaf(t0.w, jv0, fmaf(t1.x, jv1, t1.y * jv2)))));
            row[3 + a] = on ? du : 0.f;
            row[NA + 3 + a] = on ? dv : 0.f;
        }
        if (bulk_ok) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                for (int ff = 0; ff < nf; ++ff)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                                     J_out + ((size_t)(f0 + ff) * C + c) * PER_F),
                                 "r"(smem_addr(&S.stage[c & 1][ff][0])), "r"(PER_F * 4)
                                 : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            __syncthreads();
            for (int i = tid; i < nf * PER_F; i += NT) {
                const int ff = i / PER_F, r = i - ff * PER_F;
                J_out[((size_t)(f0 + ff) * C + c) * PER_F + r] = S.stage[c & 1][ff][r];
            }
            __syncthreads();
        }
    }
    if (bulk_ok && J_out && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

cudaError_t launch_fte_jac(const SceneF& scene, int n_frames, const float* x, float* uv, float* J, cudaStream_t stream) {
    if (n_frames <= 0) return cudaSuccess;
    const int grid = (n_frames + FTJ - 1) / FTJ;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fte_jac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemJ));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int bulk_ok = (((uintptr_t)J & 15u) == 0) ? 1 : 0;      // 4000-byte tiles: the base pointer decides
    fte_jac_kernel<<<grid, FTJ * NL, sizeof(SmemJ), stream>>>(scene, n_frames, bulk_ok, x, uv, J);
    return cudaGetLastError();
}

}  // namespace acino
