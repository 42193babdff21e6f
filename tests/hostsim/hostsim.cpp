// TEST HARNESS ONLY: compiles the per-env device pipeline (gym_genesis_b200/csrc/gg_env.cuh) for the
// host so that the kernel math can be checked against the oracle in the CPU-only test tier.  The
// product never loads this library; the product path is the CUDA build of the same header.
#include <stdint.h>
#include <string.h>
#include "gg_env.cuh"
#include "gg_ik.cuh"

template <int NA, int NF, bool FH>
static void run(const GgModel* M, int n, float* q, float* qd, float* qacc, const float* target, int iterations, float tol,
                int32_t* info, float* grad) {
    constexpr int NU = NA + NF, NQ = NU + 7, NV = NU + 6;
    gg::GgChain C; gg::make_chain(*M, C);
    for (int e = 0; e < n; e++) {
        gg::EnvState<NU> s;
        float* pq = q + (size_t)e * NQ; float* pv = qd + (size_t)e * NV; float* pa = qacc + (size_t)e * NV;
        for (int i = 0; i < NU; i++) { s.q[i] = pq[i]; s.qd[i] = pv[i]; }
        s.cp = gg::mk(pq[NU], pq[NU + 1], pq[NU + 2]); s.cq = gg::mkq(pq[NU + 3], pq[NU + 4], pq[NU + 5], pq[NU + 6]);
        s.cv = gg::mk(pv[NU], pv[NU + 1], pv[NU + 2]); s.cw = gg::mk(pv[NU + 3], pv[NU + 4], pv[NU + 5]);
        for (int i = 0; i < NV; i++) s.qacc[i] = pa[i];
        gg::StepInfo si;
        gg::env_substep<NA, NF, FH>(*M, C, s, target + (size_t)e * NU, iterations, tol, si);
        for (int i = 0; i < NU; i++) { pq[i] = s.q[i]; pv[i] = s.qd[i]; }
        pq[NU] = s.cp.x; pq[NU + 1] = s.cp.y; pq[NU + 2] = s.cp.z;
        pq[NU + 3] = s.cq.w; pq[NU + 4] = s.cq.x; pq[NU + 5] = s.cq.y; pq[NU + 6] = s.cq.z;
        pv[NU] = s.cv.x; pv[NU + 1] = s.cv.y; pv[NU + 2] = s.cv.z; pv[NU + 3] = s.cw.x; pv[NU + 4] = s.cw.y; pv[NU + 5] = s.cw.z;
        for (int i = 0; i < NV; i++) pa[i] = s.qacc[i];
        if (info) { info[e * 3] = si.ncon; info[e * 3 + 1] = si.iters; info[e * 3 + 2] = 0; }
        if (grad) grad[e] = si.grad;
    }
}

template <int NA, int NF, bool FH>
static void dyn(const GgModel* M, const float* q, const float* qd, const float* target, float* Mt, float* fs) {
    gg::RobotDyn<NA, NF> D;
    gg::GgChain C; gg::make_chain(*M, C);
    gg::robot_dynamics<NA, NF, FH>(*M, C, q, qd, target, D);
    constexpr int NU = NA + NF;
    for (int i = 0; i < NU; i++) { fs[i] = D.fs[i]; for (int j = 0; j < NU; j++) Mt[i * NU + j] = D.Mt[i >= j ? GG_TRI(i, j) : GG_TRI(j, i)]; }
}

template <int NA, int NF>
static void obs(const GgModel* M, int n, const float* q, float* agent, float* envs, float* reward, uint8_t* term) {
    constexpr int NU = NA + NF, NQ = NU + 7;
    for (int e = 0; e < n; e++) {
        gg::EnvState<NU> s;
        const float* pq = q + (size_t)e * NQ;
        for (int i = 0; i < NU; i++) s.q[i] = pq[i];
        s.cp = gg::mk(pq[NU], pq[NU + 1], pq[NU + 2]); s.cq = gg::mkq(pq[NU + 3], pq[NU + 4], pq[NU + 5], pq[NU + 6]);
        bool ok;
        gg::env_observe<NA, NF>(*M, s, agent + (size_t)e * (7 + NF), envs + (size_t)e * 11, reward[e], ok);
        term[e] = ok;
    }
}

extern "C" {
int hs_sizeof_model(void) { return (int)sizeof(GgModel); }
void hs_step(const GgModel* M, int n, float* q, float* qd, float* qacc, const float* target, int iterations, float tol, int32_t* info, float* grad) {
    if (M->n_arm == 7) run<7, 2, false>(M, n, q, qd, qacc, target, iterations, tol, info, grad);
    else run<5, 1, true>(M, n, q, qd, qacc, target, iterations, tol, info, grad);
}
void hs_dynamics(const GgModel* M, const float* q, const float* qd, const float* target, float* Mt, float* fs) {
    if (M->n_arm == 7) dyn<7, 2, false>(M, q, qd, target, Mt, fs); else dyn<5, 1, true>(M, q, qd, target, Mt, fs);
}
void hs_observe(const GgModel* M, int n, const float* q, float* agent, float* envs, float* reward, uint8_t* term) {
    if (M->n_arm == 7) obs<7, 2>(M, n, q, agent, envs, reward, term); else obs<5, 1>(M, n, q, agent, envs, reward, term);
}
void hs_spawn(const GgModel* M, uint64_t seed, uint32_t gid, uint32_t episode, float* xy) { gg::spawn_draw(*M, seed, gid, episode, xy[0], xy[1]); }
}

// diagnostic: the kernel's body-level broad phase ("coupled" flag) for n env-major states
extern "C" void hs_pads_may_touch(const GgModel* M, int n, const float* q, uint8_t* out) {
    for (int e = 0; e < n; e++) {
        if (M->n_arm == 7) {
            const float* pq = q + (size_t)e * 16;
            gg::RobotDyn<7, 2> D;
            gg::robot_pad_frames<7, 2, false>(*M, pq, D);
            out[e] = gg::pads_may_touch<7, 2>(*M, D, gg::mk(pq[9], pq[10], pq[11]), gg::mkq(pq[12], pq[13], pq[14], pq[15]));
        } else {
            const float* pq = q + (size_t)e * 13;
            gg::RobotDyn<5, 1> D;
            gg::robot_pad_frames<5, 1, true>(*M, pq, D);
            out[e] = gg::pads_may_touch<5, 1>(*M, D, gg::mk(pq[6], pq[7], pq[8]), gg::mkq(pq[9], pq[10], pq[11], pq[12]));
        }
    }
}

// batched IK through the device header (gg_ik.cuh), one row at a time
extern "C" void hs_ik(const GgModel* M, const GgIkConfig* K, int n, const float* q_init, const float* tpos, const float* tquat,
                      float* q_out, float* err, int32_t* iters) {
    const int nu = M->n_arm + M->n_fing;
    for (int e = 0; e < n; e++) {
        float q[GG_MAX_DOF];
        for (int i = 0; i < nu; i++) q[i] = q_init[(size_t)e * nu + i];
        GgIkConfig k = *K;
        gg::V3 tp = gg::mk(0.f, 0.f, 0.f); gg::Q4 tq = gg::mkq(1.f, 0.f, 0.f, 0.f);
        if (tpos) tp = gg::ld3(tpos + (size_t)e * 3); else k.pos_mask[0] = k.pos_mask[1] = k.pos_mask[2] = 0.f;
        if (tquat) tq = gg::ldq(tquat + (size_t)e * 4); else k.rot_mask[0] = k.rot_mask[1] = k.rot_mask[2] = 0.f;
        float pe, re;
        int it = M->n_arm == 7 ? gg::env_ik<7>(*M, k, q, tp, tq, pe, re) : gg::env_ik<5>(*M, k, q, tp, tq, pe, re);
        // re-sampling of unsolved rows, sequential form of ik_resample_kernel (same draws, same selection rule)
        float qb[GG_MAX_DOF], bp = pe, br = re, best = gg::ik_score(k, pe, re);
        for (int i = 0; i < nu; i++) qb[i] = q[i];
        bool ok = pe < k.pos_tol && re < k.rot_tol;
        for (int s = 1; s < k.max_samples && !ok; s++) {
            for (int i = 0; i < nu; i++) q[i] = qb[i];
            if (M->n_arm == 7) gg::ik_sample_init<7>(*M, k.seed_lo, k.seed_hi, (uint32_t)e, (uint32_t)s, q);
            else gg::ik_sample_init<5>(*M, k.seed_lo, k.seed_hi, (uint32_t)e, (uint32_t)s, q);
            it += M->n_arm == 7 ? gg::env_ik<7>(*M, k, q, tp, tq, pe, re) : gg::env_ik<5>(*M, k, q, tp, tq, pe, re);
            ok = pe < k.pos_tol && re < k.rot_tol;
            const float sc = gg::ik_score(k, pe, re);
            if (ok || sc < best) { best = sc; bp = pe; br = re; for (int i = 0; i < nu; i++) qb[i] = q[i]; }
        }
        for (int i = 0; i < nu; i++) q[i] = qb[i];
        pe = bp; re = br;
        for (int i = 0; i < nu; i++) q_out[(size_t)e * nu + i] = q[i];
        if (err) { err[2 * e] = pe; err[2 * e + 1] = re; }
        if (iters) iters[e] = it;
    }
}
