// Batched inverse kinematics of the end-effector link: the caller-side step every expert policy of the reference runs
// before env.step (`robot.inverse_kinematics(link=eef, pos=, quat=)`, examples/franka/pick_cube_state.py:46-51,
// examples/franka/stack_cube_state.py:78-83, examples/so_101/collect_task_stack_cube_batch.py:88-94).
//
// Algorithm (Genesis `RigidEntity.inverse_kinematics`, restated from its published behaviour, [UPSTREAM-RECALL]):
// damped least squares on the 6-dim pose error e = [target_pos - pos ; rotvec(target_quat * conj(quat))]
//     dq = J^T (J J^T + damping^2 I)^-1 e,   |dq_i| clamped to max_step_size,   q clipped to the joint limits,
// repeated until |e_pos| < pos_tol and |e_rot| < rot_tol or max_solver_iters updates have been applied.  Rows that are still
// unsolved are re-started from random arm configurations (Genesis' max_samples; its draws come from a host RNG stream, ours from
// the device Philox stream -- same distribution, uniform inside the joint limits): ik_sample_init below.  The samples of a row are
// independent, so the CUDA path runs them as the 32 LANES of one warp per unsolved row (ik_resample_kernel) and keeps the first
// success in sample order -- what a sequential loop over the samples returns.
//
// Formulation used here: with the motion subspaces S_i = (k_i, o_i x k_i) about the WORLD ORIGIN (the same vectors the
// dynamics pass uses), the Jacobian at the end-effector point p is J = X(p) S with a fixed 6x6 shift X, so
//     J J^T + l^2 I = X (G + l^2 B(p)) X^T,   G = sum_i S_i S_i^T,   B = X^-1 X^-T = [[I, -[p]x], [[p]x, (1 + |p|^2) I - p p^T]]
// and dq_i = S_i . z with (G + l^2 B) z = X^-1 e.  G is accumulated inside the chain walk (21 numbers, no per-link Jacobian
// storage) and the 6x6 system is factorised in registers by the same packed Cholesky the dynamics use.
#pragma once
#include "gg_b200.h"
#include "gg_math.cuh"

#if defined(__CUDA_ARCH__)
#define GG_IK_LOOP _Pragma("unroll 1")
#define GG_WARP_ALL(x) __all_sync(0xffffffffu, (x))
#else
#define GG_IK_LOOP
#define GG_WARP_ALL(x) (x)
#endif

namespace gg {

// rotation vector of a unit quaternion (shortest rotation)
GG_HD V3 quat_rotvec(Q4 q) {
    const float sg = q.w < 0.f ? -1.f : 1.f;
    const V3 v = mk(q.x, q.y, q.z);
    const float s = sqrtf(dot(v, v));
    float k = 2.f * sg;                                   // small angle: rotvec ~ 2 v
    if (s > 1e-9f) k = sg * 2.f * atan2f(s, fabsf(q.w)) / s;
    return v * k;
}

// start configuration of re-sampling attempt s >= 1 of row `row`: arm joints uniform inside their limits (float32 arithmetic)
template <int NA>
GG_HD void ik_sample_init(const GgModel& M, uint32_t seed_lo, uint32_t seed_hi, uint32_t row, uint32_t s, float* q) {
#pragma unroll
    for (int blk = 0; blk < (NA + 3) / 4; blk++) {
        uint32_t o[4];
        philox4x32_10(row, s, 0x494B5253u, (uint32_t)blk, seed_lo, seed_hi, o);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = blk * 4 + k;
            if (i < NA) q[i] = fmaf(u01(o[k]), M.lim_hi[i] - M.lim_lo[i], M.lim_lo[i]);
        }
    }
}

// ranking of attempts that miss the tolerances: the larger of the two tolerance-relative errors
GG_HD float ik_score(const GgIkConfig& K, float perr, float rerr) { return fmaxf(perr / K.pos_tol, rerr / K.rot_tol); }

// one env; q[0 .. NA) is updated in place (gripper coordinates do not move the end-effector link and stay untouched).
// Returns the number of DLS updates applied; perr / rerr are the error norms of the returned configuration.
template <int NA>
GG_HD int env_ik(const GgModel& M, const GgIkConfig& K, float* q, V3 tp, Q4 tq, float& perr, float& rerr) {
    const float lam2 = K.damping * K.damping;
    int it = 0, used = 0;
    for (;;) {
        // ---- chain walk: frames, motion subspaces about the world origin, G = sum S S^T (packed lower, order a.xyz l.xyz)
        SV S[NA];
        float G[21];
#pragma unroll
        for (int i = 0; i < 21; i++) G[i] = 0.f;
        Q4 ql = ldq(M.base_quat);
        V3 pl = ld3(M.base_pos);
        M3 Rl = qmat(ql);
        GG_IK_LOOP
        for (int l = 0; l < NA; l++) {
            pl = pl + mul(Rl, ld3(M.link_pos[l]));
            float s, c;
            gg_sincos(0.5f * q[l], &s, &c);
            ql = qmul_z(qmul(ql, ldq(M.link_quat[l])), c, s);
            Rl = qmat(ql);
            SV Sl; Sl.a = Rl.c2; Sl.l = cross(pl, Rl.c2);
            S[l] = Sl;
            const float v[6] = {Sl.a.x, Sl.a.y, Sl.a.z, Sl.l.x, Sl.l.y, Sl.l.z};
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) G[GG_TRI(i, j)] += v[i] * v[j];
        }
        const V3 ep = pl + mul(Rl, ld3(M.eef_pos));
        const Q4 eq = qmul(ql, ldq(M.eef_quat));
        // ---- pose error (world frame), masked per axis
        V3 dp = tp - ep;
        dp = mk(dp.x * K.pos_mask[0], dp.y * K.pos_mask[1], dp.z * K.pos_mask[2]);
        V3 dr = quat_rotvec(qmul(tq, mkq(eq.w, -eq.x, -eq.y, -eq.z)));
        dr = mk(dr.x * K.rot_mask[0], dr.y * K.rot_mask[1], dr.z * K.rot_mask[2]);
        perr = sqrtf(dot(dp, dp)); rerr = sqrtf(dot(dr, dr));
        const bool done = perr < K.pos_tol && rerr < K.rot_tol;
        if (GG_WARP_ALL(done) || it >= K.max_iters) break;
        it++;
        if (!done) {
            used++;
            // ---- (G + lam^2 B) z = X^-1 e,  X^-1 (lin, ang) = (a = ang, l = lin + p x ang)
            const float pp = dot(ep, ep);
            G[GG_TRI(0, 0)] += lam2; G[GG_TRI(1, 1)] += lam2; G[GG_TRI(2, 2)] += lam2;
            // lower-left block (rows l, cols a) = [p]x
            G[GG_TRI(3, 1)] += lam2 * -ep.z; G[GG_TRI(3, 2)] += lam2 * ep.y;
            G[GG_TRI(4, 0)] += lam2 * ep.z;  G[GG_TRI(4, 2)] += lam2 * -ep.x;
            G[GG_TRI(5, 0)] += lam2 * -ep.y; G[GG_TRI(5, 1)] += lam2 * ep.x;
            G[GG_TRI(3, 3)] += lam2 * (1.f + pp - ep.x * ep.x); G[GG_TRI(4, 4)] += lam2 * (1.f + pp - ep.y * ep.y);
            G[GG_TRI(5, 5)] += lam2 * (1.f + pp - ep.z * ep.z);
            G[GG_TRI(4, 3)] += lam2 * -ep.x * ep.y; G[GG_TRI(5, 3)] += lam2 * -ep.x * ep.z; G[GG_TRI(5, 4)] += lam2 * -ep.y * ep.z;
            const V3 rl = dp + cross(ep, dr);
            float r[6] = {dr.x, dr.y, dr.z, rl.x, rl.y, rl.z}, z[6], dinv[6];
            chol_packed<6>(G, dinv);
            chol_solve_packed<6>(G, dinv, r, z);
            SV Z; Z.a = mk(z[0], z[1], z[2]); Z.l = mk(z[3], z[4], z[5]);
            GG_IK_LOOP
            for (int l = 0; l < NA; l++) {
                float d = sdot(S[l], Z);
                d = fminf(fmaxf(d, -K.max_step), K.max_step);
                float nq = q[l] + d;
                if (K.respect_limits) nq = fminf(fmaxf(nq, M.lim_lo[l]), M.lim_hi[l]);
                q[l] = nq;
            }
        }
    }
    return used;
}

}  // namespace gg
