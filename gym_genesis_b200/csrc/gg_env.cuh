// Per-env pipeline of one env.step (one rigid substep of dt = 0.01 s), written once as
// host/device code: the CUDA kernels in gg_kernels.cu map one thread to one env and call these
// functions; tests/hostsim compiles the same header for the CPU as a debugging harness only.
//
// Replaces, for this path, what Genesis' `scene.step()` does for the reference
// (gym_genesis/tasks/franka/cube_pick.py:125; pipeline per SURVEY.md Appendix A):
//   PD torque -> FK -> CRBA (+armature, +dt*(damping+kv)) -> Cholesky -> RNEA bias -> smooth acc
//   -> narrow phase (box-plane, pad-box vs cube) -> soft-constraint Newton solve -> semi-implicit Euler.
#pragma once
#include "gg_b200.h"
#include "gg_math.cuh"

// The per-link loops of the arm chain stay ROLLED on the device (per-link temporaries then live in local
// memory): with ~1 warp per scheduler the kernel is instruction-fetch bound (ncu: stall_no_inst 67 % when fully
// unrolled), so a loop body that is fetched once and re-used from the instruction cache beats straight-line
// code (measured: 71.5 -> 48.9 us per step at 16384 envs, profiles/r01_notes.md).  -DGG_UNROLL_LINKS restores it.
#if !defined(GG_UNROLL_LINKS) && defined(__CUDA_ARCH__)
#define GG_LINK_LOOP _Pragma("unroll 1")
#else
#define GG_LINK_LOOP _Pragma("unroll")
#endif

// inner loops of the mass-matrix rows (independent iterations): -DGG_MROW_UNROLL=k trades code size for instruction-level parallelism
#if defined(GG_MROW_UNROLL) && defined(__CUDA_ARCH__)
#define GG_STR2(x) #x
#define GG_STR(x) GG_STR2(x)
#define GG_MROW_LOOP _Pragma(GG_STR(unroll GG_MROW_UNROLL))
#else
#define GG_MROW_LOOP GG_LINK_LOOP
#endif

// Same for the per-contact loops of the cube-plane solver (42.9 vs 48.9 us cold at 16384 envs).
#if !defined(GG_UNROLL_CONTACTS) && defined(__CUDA_ARCH__)
#define GG_CON_LOOP _Pragma("unroll 1")
#else
#define GG_CON_LOOP _Pragma("unroll")
#endif

// warp-uniform "does any lane ..." (callers guarantee a converged warp); the host harness runs one env at a time
#if defined(__CUDA_ARCH__) && !defined(GG_NO_LIMIT_SKIP)
#define GG_WARP_ANY(x) __any_sync(0xffffffffu, (x))
#else
#define GG_WARP_ANY(x) (x)
#endif

namespace gg {

template <int NU> struct EnvState {
    float q[NU], qd[NU];
    V3 cp; Q4 cq; V3 cv, cw;      // cube position, quaternion (wxyz), linear velocity (world), angular velocity (body frame)
    float qacc[NU + 6];           // previous solution = warm start
};

struct StepInfo { int ncon; int iters; float grad; };

// ---------------------------------------------------------------------------------------------
// Robot kinematics + dynamics.  World-frame spatial algebra about the world origin: composite
// inertias and forces of a subtree are plain sums, M(i,j) = S_i . (Ic_j S_j), c_j = S_j . Fc_j.
// ---------------------------------------------------------------------------------------------
template <int NA, int NF> struct RobotDyn {
    float Mt[(NA + NF) * (NA + NF + 1) / 2];   // M + armature + dt*(damping+kv), packed lower
    float fs[NA + NF];                         // tau - bias
    V3 jk[NA], jo[NA];                         // hinge axes / anchors (world)
    V3 fk[NF > 0 ? NF : 1], fo[NF > 0 ? NF : 1];   // finger slide direction or hinge axis, anchor
    Q4 lq; V3 lp; M3 lR;                       // last arm link frame
    M3 fR[NF > 0 ? NF : 1]; V3 fp[NF > 0 ? NF : 1];   // finger frames
};

// The dynamics pass is split in two independent halves so that the CUDA kernel can run them in different warps
// (role specialisation, gg_kernels.cu): both walk the same kinematic chain; robot_crba needs positions only.
//   robot_crba : frames, composite inertias, Mt = M + armature + dt*(damping + kv)
//   robot_rnea : velocities / accelerations / link forces, fs = clamp(PD) - damping*qd - bias
// Per-link temporaries of the rolled chain loops (motion subspace S_l, spatial inertia I_l).  The host harness and the RNEA
// warp keep them in a plain local array; the MAIN warp of the CUDA kernel passes a view of shared memory ([element][lane], conflict
// free) so that the dynamically indexed accesses are LDS / STS instead of L1-cached local memory.
template <int NA> struct LinkStoreLocal {
    SV S[NA]; SI I[NA];
    GG_HD SV getS(int l) const { return S[l]; }
    GG_HD void setS(int l, const SV& v) { S[l] = v; }
    GG_HD SI getI(int l) const { return I[l]; }
    GG_HD void setI(int l, const SI& v) { I[l] = v; }
};
template <int NA> struct LinkStoreStrided {      // b points at this lane's column; element e lives at b[e * STRIDE]
    enum { STRIDE = 32, FLOATS = NA * 16 };
    float* b;
    GG_HD SV getS(int l) const {
        const float* p = b + l * 16 * STRIDE;
        SV v; v.a = mk(p[0], p[STRIDE], p[2 * STRIDE]); v.l = mk(p[3 * STRIDE], p[4 * STRIDE], p[5 * STRIDE]);
        return v;
    }
    GG_HD void setS(int l, const SV& v) {
        float* p = b + l * 16 * STRIDE;
        p[0] = v.a.x; p[STRIDE] = v.a.y; p[2 * STRIDE] = v.a.z; p[3 * STRIDE] = v.l.x; p[4 * STRIDE] = v.l.y; p[5 * STRIDE] = v.l.z;
    }
    GG_HD SI getI(int l) const {
        const float* p = b + (l * 16 + 6) * STRIDE;
        SI v; v.m = p[0]; v.h = mk(p[STRIDE], p[2 * STRIDE], p[3 * STRIDE]);
        v.I.xx = p[4 * STRIDE]; v.I.yy = p[5 * STRIDE]; v.I.zz = p[6 * STRIDE]; v.I.xy = p[7 * STRIDE]; v.I.xz = p[8 * STRIDE]; v.I.yz = p[9 * STRIDE];
        return v;
    }
    GG_HD void setI(int l, const SI& v) {
        float* p = b + (l * 16 + 6) * STRIDE;
        p[0] = v.m; p[STRIDE] = v.h.x; p[2 * STRIDE] = v.h.y; p[3 * STRIDE] = v.h.z;
        p[4 * STRIDE] = v.I.xx; p[5 * STRIDE] = v.I.yy; p[6 * STRIDE] = v.I.zz; p[7 * STRIDE] = v.I.xy; p[8 * STRIDE] = v.I.xz; p[9 * STRIDE] = v.I.yz;
    }
};

// Chain table: the model constants of chain element l (arm link l < NA, finger l - NA) in ONE array, so that the rolled chain loops
// below read them with a single dynamically indexed constant load (the fingers are children of the last arm link and used to be
// two further unrolled copies of the loop body: ~700 SASS rows of straight-line code in a kernel that is instruction-fetch bound).
// Internal to the library (not part of the C ABI): built on the host from GgModel at launch time (make_chain).
struct GgChain { float e[GG_MAX_DOF][20]; };
enum { CH_POS = 0, CH_QUAT = 3, CH_MASS = 7, CH_COM = 8, CH_INERTIA = 11, CH_AXIS = 17 };
inline void make_chain(const GgModel& M, GgChain& C) {
    for (int l = 0; l < GG_MAX_DOF; l++) for (int k = 0; k < 20; k++) C.e[l][k] = 0.f;
    for (int l = 0; l < M.n_arm + M.n_fing && l < GG_MAX_DOF; l++) {
        float* e = C.e[l];
        const bool fin = l >= M.n_arm; const int f = l - M.n_arm;
        for (int k = 0; k < 3; k++) { e[CH_POS + k] = fin ? M.fing_pos[f][k] : M.link_pos[l][k]; e[CH_COM + k] = fin ? M.fing_com[f][k] : M.link_com[l][k]; }
        for (int k = 0; k < 4; k++) e[CH_QUAT + k] = fin ? M.fing_quat[f][k] : M.link_quat[l][k];
        for (int k = 0; k < 6; k++) e[CH_INERTIA + k] = fin ? M.fing_inertia[f][k] : M.link_inertia[l][k];
        e[CH_MASS] = fin ? M.fing_mass[f] : M.link_mass[l];
        for (int k = 0; k < 3; k++) e[CH_AXIS + k] = fin ? M.fing_axis[f][k] : (k == 2 ? 1.f : 0.f);
    }
}

// One element of the chain: parent frame (ql, pl, Rl) -> own frame, motion subspace S (world orientation, about the world origin).
// Hinges turn about their local z axis; `slide` (Panda fingers) translates along the element's axis.
GG_HD void chain_step(const float* e, bool slide, float qj, Q4& ql, V3& pl, M3& Rl, SV& S) {
    V3 p0 = pl + mul(Rl, ld3(e + CH_POS));
    Q4 qn = qmul(ql, ldq(e + CH_QUAT));
    if (!slide) {
        float s, c;
        gg_sincos(0.5f * qj, &s, &c);
        qn = qmul_z(qn, c, s);
    }
    const M3 Rn = qmat(qn);
    if (slide) {
        const V3 a = mul(Rn, ld3(e + CH_AXIS));
        S.a = mk(0.f, 0.f, 0.f); S.l = a; p0 = p0 + a * qj;
    } else {
        S.a = Rn.c2; S.l = cross(p0, Rn.c2);
    }
    ql = qn; pl = p0; Rl = Rn;
}

// finger f of a RobotDyn with a statically resolved index (keeps the small per-finger arrays in registers inside the rolled loops)
template <int NA, int NF> GG_HD void set_finger(RobotDyn<NA, NF>& D, int f, V3 k, V3 o, const M3& R, V3 p) {
#pragma unroll
    for (int g = 0; g < NF; g++) if (g == f) { D.fk[g] = k; D.fo[g] = o; D.fR[g] = R; D.fp[g] = p; }
}

template <int NA, int NF, bool FH, class LS>
GG_HD void robot_crba(const GgModel& M, const GgChain& C, const float* q, RobotDyn<NA, NF>& D, LS& ls) {
    constexpr int NU = NA + NF;
    Q4 ql = ldq(M.base_quat);
    V3 pl = ld3(M.base_pos);
    M3 Rl = qmat(ql);
    Q4 qa = ql; V3 pa = pl; M3 Ra = Rl;          // frame of the last arm link (parent of the fingers)
    SI Ic; Ic.m = 0.f; Ic.h = mk(0.f, 0.f, 0.f); Ic.I.xx = Ic.I.yy = Ic.I.zz = Ic.I.xy = Ic.I.xz = Ic.I.yz = 0.f;
    GG_LINK_LOOP
    for (int l = 0; l < NU; l++) {
        const bool fin = l >= NA;
        if (fin) { ql = qa; pl = pa; Rl = Ra; }
        const float* e = C.e[l];
        SV Sl;
        chain_step(e, fin && !FH, q[l], ql, pl, Rl, Sl);
        const SI Il = body_inertia(e[CH_MASS], ld3(e + CH_COM), e + CH_INERTIA, Rl, pl);
        if (!fin) {
            D.jk[l] = Rl.c2; D.jo[l] = pl;
            ls.setS(l, Sl);
            ls.setI(l, Il);
            if (l == NA - 1) { qa = ql; pa = pl; Ra = Rl; }
        } else {
            // leaf body: its rows of M are complete here (M(f, i) = S_i . (I_f S_f) for the arm dofs, 0 against the other finger)
            set_finger<NA, NF>(D, l - NA, FH ? Sl.a : Sl.l, pl, Rl, pl);
            const SV G = apply(Il, Sl);
            D.Mt[GG_TRI(l, l)] = sdot(Sl, G);
            GG_MROW_LOOP
            for (int i = 0; i < NA; i++) D.Mt[GG_TRI(l, i)] = sdot(ls.getS(i), G);
            for (int g = NA; g < l; g++) D.Mt[GG_TRI(l, g)] = 0.f;
            Ic = Ic + Il;
        }
    }
    D.lq = qa; D.lp = pa; D.lR = Ra;
    GG_LINK_LOOP
    for (int l = NA - 1; l >= 0; l--) {
        Ic = Ic + ls.getI(l);
        SV G = apply(Ic, ls.getS(l));
        GG_MROW_LOOP
        for (int i = 0; i <= l; i++) D.Mt[GG_TRI(l, i)] = sdot(ls.getS(i), G);
    }
    // armature; approximate implicit-fast diagonal
#pragma unroll
    for (int d = 0; d < NU; d++) D.Mt[GG_TRI(d, d)] += M.armature[d] + M.dt * (M.damping[d] + M.kv[d]);
}

// The targets are only needed by the PD law at the very end of the pass, so they come through a functor: `tf.ready()` is called
// once right before the first `tf(d)` (the CUDA kernel waits there for its asynchronously staged action slab).
// NOTE: bias[] must stay a local of this function -- hoisting it into the caller (a separate bias pass + PD pass) was
// miscompiled by nvcc 12.9 (dynamically indexed stores of the rolled loop and the static reads of the PD loop ended up at
// different stack offsets; caught by the GPU parity tests).
struct TargetPtr {
    const float* t;
    GG_HD void ready() {}
    GG_HD float operator()(int d) const { return t[d]; }
};

// Mass-matrix pass without kinematics: S and I of the arm links and the summed finger inertias come from the store the bias-force
// pass filled (DynStoreStrided; the finger rows are complete already).  Same sums in the same order as robot_crba, hence the same bits.
template <int NA, int NF, bool FH, class RS>
GG_HD void robot_crba_backward(const GgModel& M, RobotDyn<NA, NF>& D, const RS& rs) {
    constexpr int NU = NA + NF;
    SI Ic = rs.getIfingers();
    GG_LINK_LOOP
    for (int l = NA - 1; l >= 0; l--) {
        Ic = Ic + rs.getI(l);
        SV G = apply(Ic, rs.getS(l));
        GG_MROW_LOOP
        for (int i = 0; i <= l; i++) D.Mt[GG_TRI(l, i)] = sdot(rs.getS(i), G);
    }
#pragma unroll
    for (int d = 0; d < NU; d++) D.Mt[GG_TRI(d, d)] += M.armature[d] + M.dt * (M.damping[d] + M.kv[d]);
}

// per-link motion subspace / force of the bias-force pass: thread-local arrays, or this lane's column of a shared-memory scratch
template <int NA> struct RneaStoreLocal {
    enum { kFull = 0 };
    SV S[NA]; SV F[NA];
    GG_HD SV getS(int l) const { return S[l]; }
    GG_HD SV getF(int l) const { return F[l]; }
    GG_HD void set(int l, const SV& s, const SV& f) { S[l] = s; F[l] = f; }
};
template <int NA> struct RneaStoreStrided {      // element e of link l at b[(l * 12 + e) * STRIDE]
    enum { STRIDE = 32, FLOATS = NA * 12, kFull = 0 };
    float* b;
    GG_HD SV get6(const float* p) const { SV v; v.a = mk(p[0], p[STRIDE], p[2 * STRIDE]); v.l = mk(p[3 * STRIDE], p[4 * STRIDE], p[5 * STRIDE]); return v; }
    GG_HD SV getS(int l) const { return get6(b + l * 12 * STRIDE); }
    GG_HD SV getF(int l) const { return get6(b + (l * 12 + 6) * STRIDE); }
    GG_HD void set(int l, const SV& s, const SV& f) {
        float* p = b + l * 12 * STRIDE;
        p[0] = s.a.x; p[STRIDE] = s.a.y; p[2 * STRIDE] = s.a.z; p[3 * STRIDE] = s.l.x; p[4 * STRIDE] = s.l.y; p[5 * STRIDE] = s.l.z;
        p[6 * STRIDE] = f.a.x; p[7 * STRIDE] = f.a.y; p[8 * STRIDE] = f.a.z; p[9 * STRIDE] = f.l.x; p[10 * STRIDE] = f.l.y; p[11 * STRIDE] = f.l.z;
    }
};

// The thread-per-env kernel runs the bias-force pass and the mass-matrix pass in ONE thread: with this store the first pass keeps the
// motion subspace S and the world-frame spatial inertia I of every arm link in shared memory (the layout of LinkStoreStrided, + one slot
// for the summed finger inertias; the link forces F stay thread-local), fills the finger rows of the mass matrix on the way, and the
// second pass (robot_crba_backward) needs no kinematics of its own -- the chain walk and the inertia rotation are done once per step
// instead of twice (~270 instructions per link).
template <int NA> struct DynStoreStrided {
    enum { STRIDE = 32, FLOATS = NA * 16 + 10, kFull = 1 };
    float* b;
    SV F[NA];
    GG_HD SV getS(int l) const {
        const float* p = b + l * 16 * STRIDE;
        SV v; v.a = mk(p[0], p[STRIDE], p[2 * STRIDE]); v.l = mk(p[3 * STRIDE], p[4 * STRIDE], p[5 * STRIDE]);
        return v;
    }
    GG_HD SV getF(int l) const { return F[l]; }
    GG_HD void set(int l, const SV& v, const SV& f) {
        float* p = b + l * 16 * STRIDE;
        p[0] = v.a.x; p[STRIDE] = v.a.y; p[2 * STRIDE] = v.a.z; p[3 * STRIDE] = v.l.x; p[4 * STRIDE] = v.l.y; p[5 * STRIDE] = v.l.z;
        F[l] = f;
    }
    GG_HD SI getIat(const float* p) const {
        SI v; v.m = p[0]; v.h = mk(p[STRIDE], p[2 * STRIDE], p[3 * STRIDE]);
        v.I.xx = p[4 * STRIDE]; v.I.yy = p[5 * STRIDE]; v.I.zz = p[6 * STRIDE]; v.I.xy = p[7 * STRIDE]; v.I.xz = p[8 * STRIDE]; v.I.yz = p[9 * STRIDE];
        return v;
    }
    GG_HD void setIat(float* p, const SI& v) const {
        p[0] = v.m; p[STRIDE] = v.h.x; p[2 * STRIDE] = v.h.y; p[3 * STRIDE] = v.h.z;
        p[4 * STRIDE] = v.I.xx; p[5 * STRIDE] = v.I.yy; p[6 * STRIDE] = v.I.zz; p[7 * STRIDE] = v.I.xy; p[8 * STRIDE] = v.I.xz; p[9 * STRIDE] = v.I.yz;
    }
    GG_HD SI getI(int l) const { return getIat(b + (l * 16 + 6) * STRIDE); }
    GG_HD void setI(int l, const SI& v) { setIat(b + (l * 16 + 6) * STRIDE, v); }
    GG_HD SI getIfingers() const { return getIat(b + NA * 16 * STRIDE); }
    GG_HD void setIfingers(const SI& v) { setIat(b + NA * 16 * STRIDE, v); }
};

// Dp (optional): receives the frames of the pad-carrying bodies (last arm link, fingers) -- the chain walk of this pass already has
// them, so the caller's broad phase needs no walk of its own.
template <int NA, int NF, bool FH, class TF, class RS>
GG_HD void robot_rnea_t(const GgModel& M, const GgChain& C, const float* q, const float* qd, TF& tf, float* fs, RS& rs, RobotDyn<NA, NF>* Dp = nullptr) {
    constexpr int NU = NA + NF;
    float bias[NU];
    Q4 ql = ldq(M.base_quat);
    V3 pl = ld3(M.base_pos);
    M3 Rl = qmat(ql);
    SV V; V.a = mk(0.f, 0.f, 0.f); V.l = mk(0.f, 0.f, 0.f);
    SV A; A.a = mk(0.f, 0.f, 0.f); A.l = mk(0.f, 0.f, M.gravity);   // gravity as base acceleration
    Q4 qa = ql; V3 pa = pl; M3 Ra = Rl; SV Va = V, Aa = A;          // last arm link (parent of the fingers)
    SV Fc; Fc.a = mk(0.f, 0.f, 0.f); Fc.l = mk(0.f, 0.f, 0.f);
    SI Icf; Icf.m = 0.f; Icf.h = mk(0.f, 0.f, 0.f); Icf.I.xx = Icf.I.yy = Icf.I.zz = Icf.I.xy = Icf.I.xz = Icf.I.yz = 0.f;   // fingers (kFull stores)
    GG_LINK_LOOP
    for (int l = 0; l < NU; l++) {
        const bool fin = l >= NA;
        if (fin) { ql = qa; pl = pa; Rl = Ra; V = Va; A = Aa; }
        const float* e = C.e[l];
        SV Sl;
        chain_step(e, fin && !FH, q[l], ql, pl, Rl, Sl);
        A = A + crm(V, Sl) * qd[l];
        V = V + Sl * qd[l];
        const SI Il = body_inertia(e[CH_MASS], ld3(e + CH_COM), e + CH_INERTIA, Rl, pl);
        const SV Fl = apply(Il, A) + crf(V, apply(Il, V));
        if constexpr (RS::kFull != 0) {                  // everything the mass-matrix pass needs (robot_crba_backward); Dp is not null
            if (!fin) {
                rs.setI(l, Il);
                Dp->jk[l] = Rl.c2; Dp->jo[l] = pl;
            } else {
                // leaf body: its rows of M are complete here (every arm S_i is in the store already), same sums as robot_crba
                const SV G = apply(Il, Sl);
                Dp->Mt[GG_TRI(l, l)] = sdot(Sl, G);
                GG_MROW_LOOP
                for (int i = 0; i < NA; i++) Dp->Mt[GG_TRI(l, i)] = sdot(rs.getS(i), G);
                for (int g = NA; g < l; g++) Dp->Mt[GG_TRI(l, g)] = 0.f;
                Icf = Icf + Il;
            }
        }
        if (!fin) {
            rs.set(l, Sl, Fl);
            if (l == NA - 1) { qa = ql; pa = pl; Ra = Rl; Va = V; Aa = A; }
        } else {
            bias[l] = sdot(Sl, Fl);
            Fc = Fc + Fl;
            if (Dp) set_finger<NA, NF>(*Dp, l - NA, FH ? Sl.a : Sl.l, pl, Rl, pl);
        }
    }
    if (Dp) { Dp->lq = qa; Dp->lp = pa; Dp->lR = Ra; }
    if constexpr (RS::kFull != 0) rs.setIfingers(Icf);
    GG_LINK_LOOP
    for (int l = NA - 1; l >= 0; l--) {
        Fc = Fc + rs.getF(l);
        bias[l] = sdot(rs.getS(l), Fc);
    }
    // PD position control with force clamp; passive damping
    tf.ready();
#pragma unroll
    for (int d = 0; d < NU; d++) {
        float t = M.kp[d] * (tf(d) - q[d]) - M.kv[d] * qd[d];
        t = fminf(fmaxf(t, M.frc_lo[d]), M.frc_hi[d]);
        fs[d] = t - M.damping[d] * qd[d] - bias[d];
    }
}

template <int NA, int NF, bool FH>
GG_HD void robot_rnea(const GgModel& M, const GgChain& C, const float* q, const float* qd, const float* target, float* fs) {
    TargetPtr tf; tf.t = target;
    RneaStoreLocal<NA> rs;
    robot_rnea_t<NA, NF, FH>(M, C, q, qd, tf, fs, rs);
}

template <int NA, int NF, bool FH>
GG_HD void robot_dynamics(const GgModel& M, const GgChain& C, const float* q, const float* qd, const float* target, RobotDyn<NA, NF>& D) {
    LinkStoreLocal<NA> ls;
    robot_crba<NA, NF, FH>(M, C, q, D, ls);
    robot_rnea<NA, NF, FH>(M, C, q, qd, target, D.fs);
}

// frames of the pad-carrying bodies only (last arm link + fingers): the cube role's broad phase
template <int NA, int NF, bool FH>
GG_HD void robot_pad_frames(const GgModel& M, const float* q, RobotDyn<NA, NF>& D) {
    Q4 ql = ldq(M.base_quat);
    V3 pl = ld3(M.base_pos);
    M3 Rl = qmat(ql);
    GG_LINK_LOOP
    for (int l = 0; l < NA; l++) {
        pl = pl + mul(Rl, ld3(M.link_pos[l]));
        float s, c;
        gg_sincos(0.5f * q[l], &s, &c);
        ql = qmul_z(qmul(ql, ldq(M.link_quat[l])), c, s);
        Rl = qmat(ql);
    }
    D.lq = ql; D.lp = pl; D.lR = Rl;
#pragma unroll
    for (int f = 0; f < NF; f++) {
        Q4 qf = qmul(ql, ldq(M.fing_quat[f]));
        V3 p0 = pl + mul(Rl, ld3(M.fing_pos[f]));
        if (FH) {
            float s, c;
            gg_sincos(0.5f * q[NA + f], &s, &c);
            qf = qmul_z(qf, c, s);
            D.fR[f] = qmat(qf); D.fp[f] = p0;
        } else {
            M3 Rf = qmat(qf);
            D.fR[f] = Rf; D.fp[f] = p0 + mul(Rf, ld3(M.fing_axis[f])) * q[NA + f];
        }
    }
}

// pose of the end-effector frame and finger coordinates only (post-step observation refresh)
template <int NA, int NF>
GG_HD void robot_eef(const GgModel& M, const float* q, V3& ep, Q4& eq) {
    Q4 ql = ldq(M.base_quat);
    V3 pl = ld3(M.base_pos);
    M3 Rl = qmat(ql);
    GG_LINK_LOOP
    for (int l = 0; l < NA; l++) {
        pl = pl + mul(Rl, ld3(M.link_pos[l]));
        float s, c;
        gg_sincos(0.5f * q[l], &s, &c);
        ql = qmul_z(qmul(ql, ldq(M.link_quat[l])), c, s);
        Rl = qmat(ql);
    }
    ep = pl + mul(Rl, ld3(M.eef_pos));
    eq = qmul(ql, ldq(M.eef_quat));
}

// The end-effector chain is associative, so the post-step refresh is split between two warps: `robot_chain_head` walks
// links [0, KS) from the base, `robot_chain_tail` walks links [KS, NA) + the end-effector offset relative to frame KS - 1,
// `compose_eef` joins them (same result as robot_eef up to rounding).
template <int NA> GG_HD void robot_chain_head(const GgModel& M, const float* q, int ks, Q4& ql, V3& pl, M3& Rl) {
    ql = ldq(M.base_quat); pl = ld3(M.base_pos); Rl = qmat(ql);
    GG_LINK_LOOP
    for (int l = 0; l < ks; l++) {
        pl = pl + mul(Rl, ld3(M.link_pos[l]));
        float s, c;
        gg_sincos(0.5f * q[l], &s, &c);
        ql = qmul_z(qmul(ql, ldq(M.link_quat[l])), c, s);
        Rl = qmat(ql);
    }
}
// qt[l - ks] = joint coordinate of link l
template <int NA> GG_HD void robot_chain_tail(const GgModel& M, const float* qt, int ks, V3& tp, Q4& tq) {
    Q4 ql = mkq(1.f, 0.f, 0.f, 0.f);
    V3 pl = mk(0.f, 0.f, 0.f);
    M3 Rl; Rl.c0 = mk(1.f, 0.f, 0.f); Rl.c1 = mk(0.f, 1.f, 0.f); Rl.c2 = mk(0.f, 0.f, 1.f);
    GG_LINK_LOOP
    for (int l = ks; l < NA; l++) {
        pl = pl + mul(Rl, ld3(M.link_pos[l]));
        float s, c;
        gg_sincos(0.5f * qt[l - ks], &s, &c);
        ql = qmul_z(qmul(ql, ldq(M.link_quat[l])), c, s);
        Rl = qmat(ql);
    }
    tp = pl + mul(Rl, ld3(M.eef_pos));
    tq = qmul(ql, ldq(M.eef_quat));
}
GG_HD void compose_eef(Q4 ql, V3 pl, const M3& Rl, V3 tp, Q4 tq, V3& ep, Q4& eq) {
    ep = pl + mul(Rl, tp);
    eq = qmul(ql, tq);
}

// ---------------------------------------------------------------------------------------------
// Narrow phase.  Contact code: 0 = plane-cube, 1+r = plane-robot body r, 4+r = robot body r-cube
// (r: 0 = last arm link, 1+f = finger f).  Normal points from the first body to the second.
// ---------------------------------------------------------------------------------------------
template <int CAP> struct ContactsT {
    V3 p[CAP];
    V3 n[CAP];
    V3 t1[CAP];                   // first tangent (filled by the coupled solver's first pass; t2 = n x t1)
    float depth[CAP];
    int code[CAP];
    int count;
};
typedef ContactsT<GG_MAX_CONTACTS> Contacts;
typedef ContactsT<8> Contacts8;   // one geom pair (the warp-cooperative narrow phase collides one pad per lane)

template <class CT> GG_HD void add_contact(CT& C, int maxc, V3 p, V3 n, float depth, int code) {
    if (C.count < maxc) { C.p[C.count] = p; C.n[C.count] = n; C.depth[C.count] = depth; C.code[C.count] = code; C.count++; }
}

// plane z = h against a box: the first <= 4 corners (corner-index order) with signed distance <= 0
template <class CT> GG_HD void plane_box(float h, V3 c, const M3& R, V3 half, int code, int maxc, CT& C) {
    int cnt = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < 8; i++) {        // rolled: one copy of the contact append instead of eight (cold code, size matters)
        V3 l = mk((i & 1) ? half.x : -half.x, (i & 2) ? half.y : -half.y, (i & 4) ? half.z : -half.z);
        V3 w = c + mul(R, l);
        float dist = w.z - h;
        if (dist <= 0.f && cnt < 4) {
            add_contact(C, maxc, mk(w.x, w.y, w.z - 0.5f * dist), mk(0.f, 0.f, 1.f), -dist, code);
            cnt++;
        }
    }
}

GG_HD float sgn1(float x) { return x < 0.f ? -1.f : 1.f; }

// clip polygon (u, v, w) against s * coord[ax] <= lim (Sutherland-Hodgman, <= 8 vertices)
GG_HD int clip_poly(const V3* in, int n, V3* out, int ax, float s, float lim) {
    int m = 0;
    for (int i = 0; i < n; i++) {
        V3 a = in[i], b = in[(i + 1) == n ? 0 : i + 1];
        float da = s * comp(a, ax) - lim, db = s * comp(b, ax) - lim;
        if (da <= 0.f) { if (m < 8) out[m++] = a; }
        if ((da <= 0.f) != (db <= 0.f)) {
            float t = da / (da - db);
            if (m < 8) out[m++] = a + t * (b - a);
        }
    }
    return m;
}

// box A (robot pad) against box B (cube): SAT over 15 axes; A faces win ties, B faces and edge axes
// must be better by 5 %.  Face case: incident face clipped to the reference face, contacts at the
// mid-surface; edge case: midpoint of the closest points of the two supporting edges.
template <class CT> GG_HD void box_box(V3 cA, const M3& RA, V3 hA3, V3 cB, const M3& RB, V3 hB3, int code, int maxc, CT& C) {
    const float FUDGE = 1.05f;
    V3 A[3] = {RA.c0, RA.c1, RA.c2}, B[3] = {RB.c0, RB.c1, RB.c2};
    float hA[3] = {hA3.x, hA3.y, hA3.z}, hB[3] = {hB3.x, hB3.y, hB3.z};
    V3 t = cB - cA;
    float Cm[3][3], aC[3][3], tA[3], tB[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        tA[i] = dot(t, A[i]); tB[i] = dot(t, B[i]);
#pragma unroll
        for (int j = 0; j < 3; j++) { Cm[i][j] = dot(A[i], B[j]); aC[i][j] = fabsf(Cm[i][j]); }
    }
    int best = -1; float bestpen = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float pen = hA[i] + hB[0] * aC[i][0] + hB[1] * aC[i][1] + hB[2] * aC[i][2] - fabsf(tA[i]);
        if (pen < 0.f) return;
        if (best < 0 || pen < bestpen) { best = i; bestpen = pen; }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        float pen = hB[j] + hA[0] * aC[0][j] + hA[1] * aC[1][j] + hA[2] * aC[2][j] - fabsf(tB[j]);
        if (pen < 0.f) return;
        if (pen * FUDGE < bestpen) { best = 3 + j; bestpen = pen; }
    }
    V3 bestL = mk(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            V3 L = cross(A[i], B[j]);
            float l2 = dot(L, L);
            if (l2 < 1e-6f) continue;
            float l = sqrtf(l2);
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float rA = hA[i1] * aC[i2][j] + hA[i2] * aC[i1][j];
            float rB = hB[j1] * aC[i][j2] + hB[j2] * aC[i][j1];
            float pen = (rA + rB - fabsf(dot(t, L))) / l;
            if (pen < 0.f) return;
            if (pen * FUDGE < bestpen) { best = 6 + i * 3 + j; bestpen = pen; bestL = L * (1.f / l); }
        }
    }
    if (best >= 6) {
        int i = (best - 6) / 3, j = (best - 6) % 3;
        V3 n = bestL * sgn1(dot(t, bestL));
        V3 pA = cA, pB = cB;
        for (int k = 0; k < 3; k++) {
            if (k != i) pA = pA + A[k] * (sgn1(dot(n, A[k])) * hA[k]);
            if (k != j) pB = pB - B[k] * (sgn1(dot(n, B[k])) * hB[k]);
        }
        V3 d = pB - pA;
        float ab = Cm[i][j], da = dot(d, A[i]), db = dot(d, B[j]);
        float den = 1.f - ab * ab;
        float sa = (da - ab * db) / den, sb = (ab * da - db) / den;
        add_contact(C, maxc, 0.5f * (pA + A[i] * sa + pB + B[j] * sb), n, bestpen, code);
        return;
    }
    const bool refA = best < 3;
    const int ax = refA ? best : best - 3;
    V3 Rf[3], In[3]; float hR[3], hI[3];
    for (int k = 0; k < 3; k++) { Rf[k] = refA ? A[k] : B[k]; In[k] = refA ? B[k] : A[k]; hR[k] = refA ? hA[k] : hB[k]; hI[k] = refA ? hB[k] : hA[k]; }
    V3 cR = refA ? cA : cB, cI = refA ? cB : cA;
    V3 nref = Rf[ax] * sgn1(dot(cI - cR, Rf[ax]));
    int jf = 0; float bd = -1.f; float dj[3];
    for (int j = 0; j < 3; j++) { dj[j] = dot(nref, In[j]); if (fabsf(dj[j]) > bd) { bd = fabsf(dj[j]); jf = j; } }
    float fs = -sgn1(dj[jf]);
    int j1 = (jf + 1) % 3, j2 = (jf + 2) % 3, u = (ax + 1) % 3, v = (ax + 2) % 3;
    V3 poly0[8], poly1[8];
    for (int k = 0; k < 4; k++) {
        float s1 = (k == 0 || k == 3) ? 1.f : -1.f, s2 = (k < 2) ? 1.f : -1.f;
        V3 wv = cI + In[jf] * (fs * hI[jf]) + In[j1] * (s1 * hI[j1]) + In[j2] * (s2 * hI[j2]) - cR;
        poly0[k] = mk(dot(wv, Rf[u]), dot(wv, Rf[v]), dot(wv, nref));
    }
    int n = 4;
    {
        // four clips against the side planes of the reference face, ping-pong between the two buffers; ONE copy of the clipping code
        // (rolled on purpose: this path is cold code executed once per env and step, its size is what costs -- profiles/r02_notes.md)
        V3* src = poly0; V3* dst = poly1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int c = 0; c < 4; c++) {
            n = clip_poly(src, n, dst, c >> 1, (c & 1) ? -1.f : 1.f, c < 2 ? hR[u] : hR[v]);
            V3* t2 = src; src = dst; dst = t2;
        }
    }
    for (int k = 0; k < n; k++) {
        float depth = hR[ax] - poly0[k].z;
        if (depth < 0.f) continue;
        float wn = poly0[k].z + 0.5f * depth;
        V3 p = cR + Rf[u] * poly0[k].x + Rf[v] * poly0[k].y + nref * wn;
        add_contact(C, maxc, p, refA ? nref : -nref, depth, code);
    }
}

GG_HD void tangent_basis(V3 n, V3& t1, V3& t2) {
    V3 ref = fabsf(n.x) < 0.57735f ? mk(1.f, 0.f, 0.f) : mk(0.f, 1.f, 0.f);
    V3 t = ref - n * dot(ref, n);
    t1 = t * (1.f / sqrtf(dot(t, t)));
    t2 = cross(n, t1);
}

// solref / solimp impedance (power fixed at 2, checked by the host model compiler)
GG_HD float impedance(const GgModel& M, float pos) {
    float x = fabsf(pos) / M.width;
    if (x >= 1.f) return M.dmax;
    if (x <= 0.f) return M.dmin;
    float y = x <= M.mid ? x * x / M.mid : 1.f - (1.f - x) * (1.f - x) / (1.f - M.mid);
    return M.dmin + y * (M.dmax - M.dmin);
}
GG_HD void imp_aref(const GgModel& M, float pos, float vel, float& imp, float& aref) {
    imp = impedance(M, pos);
    float Bc = 2.f / (M.dmax * M.tc), Kc = 1.f / (M.dmax * M.dmax * M.tc * M.tc * M.dr * M.dr);
    aref = -Bc * vel - Kc * imp * pos;
}
GG_HD float rowval(const float* v3, float mu, int k);
// the four pyramidal rows of ONE contact share the position (hence the impedance) and the two solref constants: one impedance
// evaluation and one pair of divisions per contact instead of four (same operations in the same order per row: bit-identical)
GG_HD void imp_aref4(const GgModel& M, float pos, const float* vel3, float mu, float& imp, float* ar) {
    imp = impedance(M, pos);
    const float Bc = 2.f / (M.dmax * M.tc), Kc = 1.f / (M.dmax * M.dmax * M.tc * M.tc * M.dr * M.dr);
    const float kp = Kc * imp * pos;
#pragma unroll
    for (int k = 0; k < 4; k++) ar[k] = -Bc * rowval(vel3, mu, k) - kp;
}

template <int N> GG_HD float dotn(const float* a, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; i++) s += a[i] * b[i];
    return s;
}

// per-contact solver scalars kept between passes
struct ConRow {
    float x[GG_MAX_CONTACTS][4];     // current residual of the 4 pyramidal rows: J a - aref
    float jv[GG_MAX_CONTACTS][4];    // J s (search direction) of the same rows
    float D[GG_MAX_CONTACTS];
};

// row k of a contact: k = 0: n - mu t1, 1: n + mu t1, 2: n - mu t2, 3: n + mu t2
GG_HD float rowval(const float* v3, float mu, int k) {
    float t = k < 2 ? v3[1] : v3[2];
    return v3[0] + ((k & 1) ? mu : -mu) * t;
}

// ---------------------------------------------------------------------------------------------
// Solvers for the convex cost of A.4:  1/2 (a - a_s)^T Mt (a - a_s) + sum_active 1/2 D (J a - aref)^2,
// Newton with exact line search.  Three specialisations of the same algorithm:
//   solve_robot_limits : robot block, joint-limit rows only (J = +-e_d)          -- registers only
//   solve_cube_plane   : cube block, <= 4 plane contacts (16 pyramidal rows)      -- registers only
//   solve_coupled      : all NU + 6 dofs, any contacts (pads touching cube/plane) -- local memory
// When no pad touches anything the Hessian is block diagonal and the first two give exactly the
// minimiser of the coupled problem.  An iteration that took the full Newton step (alpha = 1) without
// any row changing its active state has landed on the exact minimiser of the current quadratic
// piece, so the loop stops there without re-assembling the gradient.
// ---------------------------------------------------------------------------------------------
// Stop criterion shared by the three solvers: largest diagonally-preconditioned gradient component
// (an estimate of the remaining acceleration error per dof) relative to 1 + |a|_inf.  Unlike a gradient
// norm scaled by the mean inertia of the whole system it stays meaningful for the cube's rotational dofs,
// whose inertia (3.4e-6 kg m^2) is six orders below the arm's.
// Second stop test of the Newton loops: the step just taken changed no acceleration by more than tol (1 + |a|_inf).  Without it a
// solve whose gradient test cannot pass in float32 (|g_i| / H_ii with H_ii ~ 1e-5 for the rotation of a softly supported cube: the
// SO-101 task's friction 5 makes the pyramidal regulariser 650x softer) keeps taking noise-sized steps that flip rows sitting at
// x = 0 until the iteration cap -- and a warp runs as long as its slowest lane (9 % of the lanes at the cap = every warp at the cap).
template <int N> GG_HD bool newton_step_small(const float* sd, float al, const float* a, float tol) {
    float sm = 0.f, am = 0.f;
#pragma unroll
    for (int i = 0; i < N; i++) { sm = fmaxf(sm, fabsf(al * sd[i])); am = fmaxf(am, fabsf(a[i])); }
    return sm <= tol * (1.f + am);
}
// Float32 noise floor of a gradient component: g_i = (M (a - a_smooth))_i + sum_rows J_row,i f_row is a sum of terms that cancel at
// the optimum, so |g_i| cannot be resolved below a few ulps of the LARGEST term; GG_GRAD_ULPS ulps of sum |term| count as zero.
// (A cube resting flat: four corner forces of 0.25 N at 28 mm lever arms cancel to a torque of 1e-10 .. 5e-10 N m = 1 .. 5 ulps of
// 7e-3 N m, which the rotational H_ii ~ 6e-6 of a softly supported cube -- the SO-101 task's friction 5 makes the pyramidal
// regulariser 650x softer -- turns into 2e-5 .. 8e-5 > tol: the gradient test could never pass and every step ran Newton
// iterations and nine-step line searches on noise, 9 % of the lanes up to the iteration cap.)
#ifndef GG_GRAD_ULPS
#define GG_GRAD_ULPS 8.f
#endif
GG_HD float grad_above_noise(float g, float sum_abs_terms) { return fmaxf(fabsf(g) - GG_GRAD_ULPS * 1.1920929e-7f * sum_abs_terms, 0.f); }
template <int N> GG_HD float newton_residual(const float* g, const float* H, const float* a) {
    float r = 0.f, am = 0.f;
#pragma unroll
    for (int i = 0; i < N; i++) { r = fmaxf(r, fabsf(g[i]) * gg_rcp(H[GG_TRI(i, i)])); am = fmaxf(am, fabsf(a[i])); }
    return gg_div(r, 1.f + am);
}

// Line-search accuracy only shapes the path of the outer Newton loop, not its fixed point (the outer loop stops on the
// residual or on an exact full step).  1e-2 relative on phi' is the upstream ls_tolerance.
#define GG_LS_TOL 1e-2f
#define GG_LS_MAX 8

struct LineSearch {
    float al, lo, hi, d0, d1, d2;
    int ls;
    GG_HD void begin() { al = 0.f; lo = 0.f; hi = -1.f; d0 = 0.f; d1 = 0.f; d2 = 1.f; ls = 0; }
    // call with phi'(al), phi''(al); returns false when converged
    GG_HD bool next(float nd1, float nd2) {
        d1 = nd1; d2 = nd2;
        if (ls == 0) d0 = d1;
        else {
            if (d1 < 0.f) lo = al; else hi = al;
            if (fabsf(d1) <= fabsf(d0) * GG_LS_TOL) return false;
        }
        if (++ls > GG_LS_MAX) return false;
        float nx = al - gg_div(d1, d2);
        if (hi >= 0.f && (nx <= lo || nx >= hi)) nx = 0.5f * (lo + hi);
        if (nx <= 0.f) nx = hi >= 0.f ? 0.5f * (lo + hi) : 1e-3f;
        al = nx;
        return true;
    }
};

template <int NU>
GG_HD void solve_robot_limits(const float* Mt, const float* as, const float* lsg, const float* laref, const float* lD,
                              const float* warm, float* a, int iterations, float tol, float scale, int& iters, float& gnorm) {
    float e[NU], Me[NU];
    float costw = 0.f, costs = 0.f;
#pragma unroll
    for (int i = 0; i < NU; i++) e[i] = warm[i] - as[i];
    symv_packed<NU>(Mt, e, Me);
#pragma unroll
    for (int d = 0; d < NU; d++) {
        costw += 0.5f * e[d] * Me[d];
        float jw = lsg[d] * warm[d] - laref[d], js = lsg[d] * as[d] - laref[d];
        if (lsg[d] != 0.f && jw < 0.f) costw += 0.5f * lD[d] * jw * jw;
        if (lsg[d] != 0.f && js < 0.f) costs += 0.5f * lD[d] * js * js;
    }
    const bool usew = costw < costs;
#pragma unroll
    for (int i = 0; i < NU; i++) a[i] = usew ? warm[i] : as[i];
    int it = 0;
    for (; it < iterations; it++) {
        float g[NU], H[NU * (NU + 1) / 2], dinv[NU], sd[NU], Ms[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) e[i] = a[i] - as[i];
        symv_packed<NU>(Mt, e, Me);
#pragma unroll
        for (int i = 0; i < NU * (NU + 1) / 2; i++) H[i] = Mt[i];
#pragma unroll
        for (int d = 0; d < NU; d++) {
            float jar = lsg[d] * a[d] - laref[d];
            bool act = lsg[d] != 0.f && jar < 0.f;
            g[d] = Me[d] + (act ? lD[d] * jar * lsg[d] : 0.f);
            H[GG_TRI(d, d)] += act ? lD[d] : 0.f;
        }
        gnorm = newton_residual<NU>(g, H, a);
        if (gnorm < tol) break;
        chol_packed<NU>(H, dinv);
#pragma unroll
        for (int i = 0; i < NU; i++) g[i] = -g[i];
        chol_solve_packed<NU>(H, dinv, g, sd);
        symv_packed<NU>(Mt, sd, Ms);
        const float q1 = dotn<NU>(sd, Me), q2 = dotn<NU>(sd, Ms);
        LineSearch L; L.begin();
        for (;;) {
            float d1 = q1 + L.al * q2, d2 = q2;
#pragma unroll
            for (int d = 0; d < NU; d++) {
                float jv = lsg[d] * sd[d];
                float x = lsg[d] * a[d] - laref[d] + L.al * jv;
                if (lsg[d] != 0.f && x < 0.f) { d1 += lD[d] * x * jv; d2 += lD[d] * jv * jv; }
            }
            if (!L.next(d1, d2)) break;
        }
        int changed = 0;
#pragma unroll
        for (int d = 0; d < NU; d++) {
            float j0 = lsg[d] * a[d] - laref[d];
            a[d] += L.al * sd[d];
            float j1 = lsg[d] * a[d] - laref[d];
            changed += (lsg[d] != 0.f) && ((j0 < 0.f) != (j1 < 0.f));
        }
        if (changed == 0 && fabsf(L.al - 1.f) < 1e-3f) { it++; break; }
    }
    iters = it;
}

// Per-contact temporaries of the cube-plane solver: 29 floats per contact -- J (3 x 6), aref (4), D, u (3), w (3).
// Local arrays on the host / shared-memory view in the CUDA kernel's CUBE warp (see LinkStoreStrided).
struct CubeConLocal {
    float v[4][29];
    GG_HD float get(int c, int k) const { return v[c][k]; }
    GG_HD void set(int c, int k, float x) { v[c][k] = x; }
};
struct CubeConStrided {
    enum { STRIDE = 32, FLOATS = 4 * 29 };
    float* b;
    GG_HD float get(int c, int k) const { return b[(c * 29 + k) * STRIDE]; }
    GG_HD void set(int c, int k, float x) { b[(c * 29 + k) * STRIDE] = x; }
};
enum { CC_J = 0, CC_AREF = 18, CC_D = 22, CC_U = 23, CC_W = 26 };

// cube block against the plane: contacts i < ncp are live, the rest are masked with D = 0
template <class CS>
GG_HD void solve_cube_plane(const GgModel& M, const float* Mc, const float* as, const float* qd, V3 cp, const M3& cR, const V3* p,
                            const float* depth, int ncp, const float* warm, float* a, int iterations, float tol, CS& cs,
                            int& iters, float& gnorm) {
    const float mu = M.friction;
    float costw = 0.f, costs = 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) { float ew = warm[i] - as[i]; costw += 0.5f * Mc[i] * ew * ew; }
    GG_CON_LOOP
    for (int c = 0; c < 4; c++) {
        // plane normal +z: n = z, t1 = x, t2 = y (= tangent_basis(z)); rows act on [v, w_body]
        V3 r = p[c] - cp;
        V3 a0 = cross(cR.c0, r), a1 = cross(cR.c1, r), a2 = cross(cR.c2, r);
        float J[3][6];
        J[0][0] = 0.f; J[0][1] = 0.f; J[0][2] = 1.f; J[0][3] = a0.z; J[0][4] = a1.z; J[0][5] = a2.z;
        J[1][0] = 1.f; J[1][1] = 0.f; J[1][2] = 0.f; J[1][3] = a0.x; J[1][4] = a1.x; J[1][5] = a2.x;
        J[2][0] = 0.f; J[2][1] = 1.f; J[2][2] = 0.f; J[2][3] = a0.y; J[2][4] = a1.y; J[2][5] = a2.y;
        float vel[3], uw[3], us[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { vel[k] = dotn<6>(J[k], qd); uw[k] = dotn<6>(J[k], warm); us[k] = dotn<6>(J[k], as); }
        const bool live = c < ncp;
        float imp = 1.f, ar[4];
        imp_aref4(M, -depth[c], vel, mu, imp, ar);
        const float Dk = live ? gg_rcp(fmaxf(M.minval, M.reg_scale * (1.f - imp) / imp * M.cube_invweight)) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float jw = rowval(uw, mu, k) - ar[k], js = rowval(us, mu, k) - ar[k];
            if (jw < 0.f) costw += 0.5f * Dk * jw * jw;
            if (js < 0.f) costs += 0.5f * Dk * js * js;
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int i = 0; i < 6; i++) cs.set(c, CC_J + k * 6 + i, J[k][i]);
#pragma unroll
        for (int k = 0; k < 4; k++) cs.set(c, CC_AREF + k, ar[k]);
        cs.set(c, CC_D, Dk);
#pragma unroll
        for (int k = 0; k < 3; k++) { cs.set(c, CC_U + k, uw[k]); cs.set(c, CC_W + k, us[k]); }
    }
    const bool usew = costw < costs;
#pragma unroll
    for (int i = 0; i < 6; i++) a[i] = usew ? warm[i] : as[i];
    if (!usew) {
        GG_CON_LOOP
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int k = 0; k < 3; k++) cs.set(c, CC_U + k, cs.get(c, CC_W + k));
    }
    int it = 0;
    for (; it < iterations; it++) {
        float H[21], g[6], gabs[6], dinv[6], sd[6];
#pragma unroll
        for (int i = 0; i < 21; i++) H[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) { H[GG_TRI(i, i)] = Mc[i]; g[i] = Mc[i] * (a[i] - as[i]); gabs[i] = fabsf(g[i]); }
        GG_CON_LOOP
        for (int c = 0; c < 4; c++) {
            float J[3][6], u[3], ar[4];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int i = 0; i < 6; i++) J[k][i] = cs.get(c, CC_J + k * 6 + i);
#pragma unroll
            for (int k = 0; k < 3; k++) u[k] = cs.get(c, CC_U + k);
#pragma unroll
            for (int k = 0; k < 4; k++) ar[k] = cs.get(c, CC_AREF + k);
            const float Dk = cs.get(c, CC_D);
            float Wnn = 0.f, Wn1 = 0.f, W11 = 0.f, Wn2 = 0.f, W22 = 0.f, gn = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float jar = rowval(u, mu, k) - ar[k];
                float on = jar < 0.f ? Dk : 0.f;
                float sm = (k & 1) ? mu : -mu;
                Wnn += on; gn += on * jar;
                if (k < 2) { Wn1 += on * sm; W11 += on * mu * mu; g1 += on * jar * sm; }
                else { Wn2 += on * sm; W22 += on * mu * mu; g2 += on * jar * sm; }
            }
            float Pn[6], P1[6], P2[6];
#pragma unroll
            for (int i = 0; i < 6; i++) {
                g[i] += J[0][i] * gn + J[1][i] * g1 + J[2][i] * g2;
                gabs[i] += fabsf(J[0][i] * gn) + fabsf(J[1][i] * g1) + fabsf(J[2][i] * g2);
                Pn[i] = Wnn * J[0][i] + Wn1 * J[1][i] + Wn2 * J[2][i];
                P1[i] = Wn1 * J[0][i] + W11 * J[1][i];
                P2[i] = Wn2 * J[0][i] + W22 * J[2][i];
            }
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) H[GG_TRI(i, j)] += J[0][i] * Pn[j] + J[1][i] * P1[j] + J[2][i] * P2[j];
        }
        {
            // stop test on the gradient above its float32 noise floor
            float r = 0.f, am = 0.f;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                r = fmaxf(r, grad_above_noise(g[i], gabs[i]) * gg_rcp(H[GG_TRI(i, i)]));
                am = fmaxf(am, fabsf(a[i]));
            }
            gnorm = gg_div(r, 1.f + am);
        }
        if (gnorm < tol) break;
        chol_packed<6>(H, dinv);
#pragma unroll
        for (int i = 0; i < 6; i++) g[i] = -g[i];
        chol_solve_packed<6>(H, dinv, g, sd);
        float q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) { q1 += Mc[i] * sd[i] * (a[i] - as[i]); q2 += Mc[i] * sd[i] * sd[i]; }
        GG_CON_LOOP
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float r = 0.f;
#pragma unroll
                for (int i = 0; i < 6; i++) r += cs.get(c, CC_J + k * 6 + i) * sd[i];
                cs.set(c, CC_W + k, r);
            }
        LineSearch L; L.begin();
        for (;;) {
            float d1 = q1 + L.al * q2, d2 = q2;
            GG_CON_LOOP
            for (int c = 0; c < 4; c++) {
                const float u[3] = {cs.get(c, CC_U), cs.get(c, CC_U + 1), cs.get(c, CC_U + 2)};
                const float w[3] = {cs.get(c, CC_W), cs.get(c, CC_W + 1), cs.get(c, CC_W + 2)};
                const float Dk = cs.get(c, CC_D);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float jv = rowval(w, mu, k);
                    float x = rowval(u, mu, k) - cs.get(c, CC_AREF + k) + L.al * jv;
                    if (x < 0.f) { d1 += Dk * x * jv; d2 += Dk * jv * jv; }
                }
            }
            if (!L.next(d1, d2)) break;
        }
        int changed = 0;
        GG_CON_LOOP
        for (int c = 0; c < 4; c++) {
            const float u[3] = {cs.get(c, CC_U), cs.get(c, CC_U + 1), cs.get(c, CC_U + 2)};
            float un[3] = {u[0] + L.al * cs.get(c, CC_W), u[1] + L.al * cs.get(c, CC_W + 1), u[2] + L.al * cs.get(c, CC_W + 2)};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float ar = cs.get(c, CC_AREF + k);
                changed += (c < ncp) && ((rowval(u, mu, k) - ar < 0.f) != (rowval(un, mu, k) - ar < 0.f));
            }
            cs.set(c, CC_U, un[0]); cs.set(c, CC_U + 1, un[1]); cs.set(c, CC_U + 2, un[2]);
        }
#pragma unroll
        for (int i = 0; i < 6; i++) a[i] += L.al * sd[i];
        if ((changed == 0 && fabsf(L.al - 1.f) < 1e-3f) || newton_step_small<6>(sd, L.al, a, tol)) { it++; break; }
    }
    iters = it;
}

// General coupled solve over all NV = NU + 6 dofs (pads in contact).
// Every contact row is a unit force d at point p, i.e. the spatial force F = (p x d, d) about a reference point O,
// and its Jacobian is  J a = F . (T_b(a) - T_a(a))  with T_x(a) = sum_i B_i a_i the spatial acceleration of body
// x (B_i = twist of dof i about O: the S_i of the dynamics pass re-referenced, plus the cube's six).  All contacts
// between the same pair of bodies share that twist basis, so the Hessian / gradient of a pair are first
// accumulated in 6-dim spatial-force space (Phi = sum D F F^T, 21 numbers; gF = sum D jar F) at ~180 instructions per
// contact and projected onto the dofs ONCE per pair (H += P^T Phi P), instead of building a 3 x 15 Jacobian per
// contact (~1500 instructions).  Pair index = contact code (0 plane-cube, 1+r plane-robot body r, 4+r body r-cube).
GG_HD SV sym6_mul(const float* P, const SV& v) {     // packed lower 6x6, order (a.x a.y a.z l.x l.y l.z)
    const float x[6] = {v.a.x, v.a.y, v.a.z, v.l.x, v.l.y, v.l.z};
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < 6; j++) r += P[i >= j ? GG_TRI(i, j) : GG_TRI(j, i)] * x[j];
        y[i] = r;
    }
    SV o; o.a = mk(y[0], y[1], y[2]); o.l = mk(y[3], y[4], y[5]);
    return o;
}

template <int NA, int NF, bool FH>
GG_HD void solve_coupled(const GgModel& M, const RobotDyn<NA, NF>& D, const EnvState<NA + NF>& s, const M3& cR, const float* Mc,
                         const float* as, const float* qdv, const float* lsg, const float* laref, const float* lD,
                         Contacts& C, float* a, int iterations, float tol, float scale, int& iters, float& gnorm_out) {
    constexpr int NU = NA + NF, NV = NU + 6, NH = NV * (NV + 1) / 2, NP = 7;
    ConRow R;
    const float mu = M.friction;
    // ---- twist basis of every dof about the reference point O = cube centre (all contacts are within a few cm of the
    //      cube or of the fingertips next to it, so lever arms stay small and fp32 cancellation is avoided)
    const V3 O = s.cp;
    SV B[NV];
#pragma unroll
    for (int i = 0; i < NA; i++) { B[i].a = D.jk[i]; B[i].l = cross(D.jo[i] - O, D.jk[i]); }
#pragma unroll
    for (int f = 0; f < NF; f++) {
        if (FH) { B[NA + f].a = D.fk[f]; B[NA + f].l = cross(D.fo[f] - O, D.fk[f]); }
        else { B[NA + f].a = mk(0.f, 0.f, 0.f); B[NA + f].l = D.fk[f]; }
    }
    B[NU + 0].a = mk(0.f, 0.f, 0.f); B[NU + 0].l = mk(1.f, 0.f, 0.f);
    B[NU + 1].a = mk(0.f, 0.f, 0.f); B[NU + 1].l = mk(0.f, 1.f, 0.f);
    B[NU + 2].a = mk(0.f, 0.f, 0.f); B[NU + 2].l = mk(0.f, 0.f, 1.f);
    B[NU + 3].a = cR.c0; B[NU + 3].l = mk(0.f, 0.f, 0.f);
    B[NU + 4].a = cR.c1; B[NU + 4].l = mk(0.f, 0.f, 0.f);
    B[NU + 5].a = cR.c2; B[NU + 5].l = mk(0.f, 0.f, 0.f);
    // relative spatial acceleration of each pair for a dof vector x
    auto pair_twists = [&](const float* x, SV* T) {
        SV arm; arm.a = mk(0.f, 0.f, 0.f); arm.l = mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NA; i++) arm = arm + B[i] * x[i];
        SV cube; cube.a = mk(0.f, 0.f, 0.f); cube.l = mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 6; j++) cube = cube + B[NU + j] * x[NU + j];
        T[0] = cube; T[1] = arm; T[4] = cube + arm * -1.f;
#pragma unroll
        for (int f = 0; f < 2; f++) {
            SV rb = arm;
            if (f < NF) rb = arm + B[NA + (f < NF ? f : 0)] * x[NA + (f < NF ? f : 0)];
            T[2 + f] = rb; T[5 + f] = cube + rb * -1.f;
        }
    };
    SV Tq[NP], Tw[NP], Ts[NP];
    pair_twists(qdv, Tq); pair_twists(s.qacc, Tw); pair_twists(as, Ts);

    float costw = 0.f, costs = 0.f;
    {
        float e[NV], Me[NU];
#pragma unroll
        for (int i = 0; i < NV; i++) e[i] = s.qacc[i] - as[i];
        symv_packed<NU>(D.Mt, e, Me);
#pragma unroll
        for (int i = 0; i < NU; i++) costw += 0.5f * e[i] * Me[i];
#pragma unroll
        for (int i = 0; i < 6; i++) costw += 0.5f * Mc[i] * e[NU + i] * e[NU + i];
#pragma unroll
        for (int d = 0; d < NU; d++) {
            float jw = lsg[d] * s.qacc[d] - laref[d], js = lsg[d] * as[d] - laref[d];
            if (lsg[d] != 0.f && jw < 0.f) costw += 0.5f * lD[d] * jw * jw;
            if (lsg[d] != 0.f && js < 0.f) costs += 0.5f * lD[d] * js * js;
        }
    }
    unsigned used = 0u;
    for (int c = 0; c < C.count; c++) {
        V3 t1, t2;
        const V3 n = C.n[c], p = C.p[c] - O;
        tangent_basis(n, t1, t2);
        C.t1[c] = t1;
        SV Fn, F1, F2;
        Fn.a = cross(p, n); Fn.l = n; F1.a = cross(p, t1); F1.l = t1; F2.a = cross(p, t2); F2.l = t2;
        const int code = C.code[c];
        used |= 1u << code;
        float vel[3] = {sdot(Fn, Tq[code]), sdot(F1, Tq[code]), sdot(F2, Tq[code])};
        float uw[3] = {sdot(Fn, Tw[code]), sdot(F1, Tw[code]), sdot(F2, Tw[code])};
        float us[3] = {sdot(Fn, Ts[code]), sdot(F1, Ts[code]), sdot(F2, Ts[code])};
        const int rb = code == 0 ? -1 : (code - 1) % 3;
        const float wsum = (rb >= 0 ? M.body_invweight[rb] : 0.f) + ((code == 0 || code >= 4) ? M.cube_invweight : 0.f);
        float imp = 1.f, ar[4];
        imp_aref4(M, -C.depth[c], vel, mu, imp, ar);
        const float Dk = gg_rcp(fmaxf(M.minval, M.reg_scale * (1.f - imp) / imp * wsum));
        R.D[c] = Dk;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float jw = rowval(uw, mu, k) - ar[k], js = rowval(us, mu, k) - ar[k];
            if (jw < 0.f) costw += 0.5f * Dk * jw * jw;
            if (js < 0.f) costs += 0.5f * Dk * js * js;
            R.x[c][k] = jw; R.jv[c][k] = js;          // both candidates; the loser is dropped below
        }
    }
    const bool warm = costw < costs;
#pragma unroll
    for (int i = 0; i < NV; i++) a[i] = warm ? s.qacc[i] : as[i];
    if (!warm) for (int c = 0; c < C.count; c++) {
#pragma unroll
        for (int k = 0; k < 4; k++) R.x[c][k] = R.jv[c][k];
    }

    int it = 0; float gnorm = 0.f;
    for (; it < iterations; it++) {
        float H[NH], g[NV], dinv[NV];
        float Phi[NP][21], gF[NP][6];
#pragma unroll
        for (int i = 0; i < NH; i++) H[i] = 0.f;
#pragma unroll
        for (int i = 0; i < NU * (NU + 1) / 2; i++) H[i] = D.Mt[i];
#pragma unroll
        for (int i = 0; i < 6; i++) H[GG_TRI(NU + i, NU + i)] = Mc[i];
        {
            float e[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) e[i] = a[i] - as[i];
            symv_packed<NU>(D.Mt, e, g);
#pragma unroll
            for (int i = 0; i < 6; i++) g[NU + i] = Mc[i] * e[NU + i];
        }
#pragma unroll
        for (int d = 0; d < NU; d++) {
            float jar = lsg[d] * a[d] - laref[d];
            if (lsg[d] != 0.f && jar < 0.f) { H[GG_TRI(d, d)] += lD[d]; g[d] += lD[d] * jar * lsg[d]; }
        }
        for (int pc = 0; pc < NP; pc++) {
            for (int i = 0; i < 21; i++) Phi[pc][i] = 0.f;
            for (int i = 0; i < 6; i++) gF[pc][i] = 0.f;
        }
        // ---- accumulate every pair's stiffness / gradient in spatial-force space
        for (int c = 0; c < C.count; c++) {
            float Wnn = 0.f, Wn1 = 0.f, W11 = 0.f, Wn2 = 0.f, W22 = 0.f, gn = 0.f, g1 = 0.f, g2 = 0.f;
            const float Dk = R.D[c];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float jar = R.x[c][k];
                if (jar < 0.f) {
                    float sm = (k & 1) ? mu : -mu;
                    Wnn += Dk; gn += Dk * jar;
                    if (k < 2) { Wn1 += Dk * sm; W11 += Dk * mu * mu; g1 += Dk * jar * sm; }
                    else { Wn2 += Dk * sm; W22 += Dk * mu * mu; g2 += Dk * jar * sm; }
                }
            }
            if (Wnn == 0.f) continue;
            const V3 n = C.n[c], p = C.p[c] - O;
            const V3 t1 = C.t1[c], t2 = cross(n, t1);
            const V3 pn = cross(p, n), p1 = cross(p, t1), p2 = cross(p, t2);
            const float fn[6] = {pn.x, pn.y, pn.z, n.x, n.y, n.z}, f1[6] = {p1.x, p1.y, p1.z, t1.x, t1.y, t1.z},
                        f2[6] = {p2.x, p2.y, p2.z, t2.x, t2.y, t2.z};
            float Pn[6], P1[6], P2[6];
            const int code = C.code[c];
#pragma unroll
            for (int i = 0; i < 6; i++) {
                gF[code][i] += fn[i] * gn + f1[i] * g1 + f2[i] * g2;
                Pn[i] = Wnn * fn[i] + Wn1 * f1[i] + Wn2 * f2[i];
                P1[i] = Wn1 * fn[i] + W11 * f1[i];
                P2[i] = Wn2 * fn[i] + W22 * f2[i];
            }
#pragma unroll
            for (int i = 0; i < 6; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) Phi[code][GG_TRI(i, j)] += fn[i] * Pn[j] + f1[i] * P1[j] + f2[i] * P2[j];
        }
        // ---- project each used pair onto its dofs: H += P^T Phi P, g += P^T gF
        for (int pc = 0; pc < NP; pc++) {
            if (!((used >> pc) & 1u)) continue;
            const int rb = pc == 0 ? -1 : (pc - 1) % 3;
            const float sr = pc >= 4 ? -1.f : 1.f;
            int idx[NV]; float sg[NV]; int nd = 0;
            if (rb >= 0) {
                for (int i = 0; i < NA; i++) { idx[nd] = i; sg[nd] = sr; nd++; }
                if (rb >= 1 && rb - 1 < NF) { idx[nd] = NA + rb - 1; sg[nd] = sr; nd++; }
            }
            if (pc == 0 || pc >= 4) for (int j = 0; j < 6; j++) { idx[nd] = NU + j; sg[nd] = 1.f; nd++; }
            SV gf; gf.a = mk(gF[pc][0], gF[pc][1], gF[pc][2]); gf.l = mk(gF[pc][3], gF[pc][4], gF[pc][5]);
            for (int jj = 0; jj < nd; jj++) {
                const int j = idx[jj];
                const SV y = sym6_mul(Phi[pc], B[j]) * sg[jj];
                g[j] += sg[jj] * sdot(B[j], gf);
                for (int ii = jj; ii < nd; ii++) H[GG_TRI(idx[ii], j)] += sg[ii] * sdot(B[idx[ii]], y);
            }
        }
        gnorm = newton_residual<NV>(g, H, a);
        if (gnorm < tol) break;
        if (!chol_packed<NV>(H, dinv)) break;
        float sdir[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) g[i] = -g[i];
        chol_solve_packed<NV>(H, dinv, g, sdir);
        float q1 = 0.f, q2 = 0.f;
        {
            float Ms[NU];
            symv_packed<NU>(D.Mt, sdir, Ms);
#pragma unroll
            for (int i = 0; i < NU; i++) { q1 += Ms[i] * (a[i] - as[i]); q2 += Ms[i] * sdir[i]; }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) { q1 += Mc[i] * sdir[NU + i] * (a[NU + i] - as[NU + i]); q2 += Mc[i] * sdir[NU + i] * sdir[NU + i]; }
        pair_twists(sdir, Ts);
        for (int c = 0; c < C.count; c++) {
            const V3 n = C.n[c], p = C.p[c] - O;
            const V3 t1 = C.t1[c], t2 = cross(n, t1);
            SV Fn, F1, F2;
            Fn.a = cross(p, n); Fn.l = n; F1.a = cross(p, t1); F1.l = t1; F2.a = cross(p, t2); F2.l = t2;
            const int code = C.code[c];
            const float w3[3] = {sdot(Fn, Ts[code]), sdot(F1, Ts[code]), sdot(F2, Ts[code])};
#pragma unroll
            for (int k = 0; k < 4; k++) R.jv[c][k] = rowval(w3, mu, k);
        }
        LineSearch L; L.begin();
        for (;;) {
            float d1 = q1 + L.al * q2, d2 = q2;
#pragma unroll
            for (int d = 0; d < NU; d++) {
                float jv = lsg[d] * sdir[d];
                float x = lsg[d] * a[d] - laref[d] + L.al * jv;
                if (lsg[d] != 0.f && x < 0.f) { d1 += lD[d] * x * jv; d2 += lD[d] * jv * jv; }
            }
            for (int c = 0; c < C.count; c++) {
                const float Dk = R.D[c];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float jv = R.jv[c][k];
                    const float x = R.x[c][k] + L.al * jv;
                    if (x < 0.f) { d1 += Dk * x * jv; d2 += Dk * jv * jv; }
                }
            }
            if (!L.next(d1, d2)) break;
        }
        const float al = L.al;
        int changed = 0;
#pragma unroll
        for (int d = 0; d < NU; d++) {
            float j0 = lsg[d] * a[d] - laref[d], j1 = lsg[d] * (a[d] + al * sdir[d]) - laref[d];
            changed += (lsg[d] != 0.f) && ((j0 < 0.f) != (j1 < 0.f));
        }
#pragma unroll
        for (int i = 0; i < NV; i++) a[i] += al * sdir[i];
        for (int c = 0; c < C.count; c++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float x0 = R.x[c][k], x1 = x0 + al * R.jv[c][k];
                changed += (x0 < 0.f) != (x1 < 0.f);
                R.x[c][k] = x1;
            }
        }
        if (changed == 0 && fabsf(al - 1.f) < 1e-3f) { it++; break; }
    }
    iters = it; gnorm_out = gnorm;
}

// pads against plane and cube in the oracle's order (after the cube-plane contacts already in C)
template <int NA, int NF>
GG_HD void collide_pads(const GgModel& M, const RobotDyn<NA, NF>& D, V3 cp, const M3& cR, Contacts& C) {
    const int maxc = M.max_contacts < GG_MAX_CONTACTS ? M.max_contacts : GG_MAX_CONTACTS;
    V3 ch = ld3(M.cube_half);
    for (int pass = 0; pass < 2; pass++) {
        for (int g = 0; g < M.n_pads; g++) {
            int b = M.pad_body[g];
            int fi = b - 1 < 0 ? 0 : (b - 1 < NF ? b - 1 : 0);
            M3 Rb = b == 0 ? D.lR : D.fR[fi];
            V3 pb = b == 0 ? D.lp : D.fp[fi];
            V3 c = pb + mul(Rb, ld3(M.pad_pos[g]));
            M3 R = mmul(Rb, qmat(ldq(M.pad_quat[g])));
            V3 h = ld3(M.pad_half[g]);
            if (pass == 0) plane_box(M.plane_z, c, R, h, 1 + b, maxc, C);
            else box_box(c, R, h, cp, cR, ch, 4 + b, maxc, C);
        }
    }
}

// exact box-plane reach of the pads of body b (lowest point of an oriented box: c.z - sum_i |R_zi| h_i), with a 1e-5 m margin so
// that it stays conservative against the corner-by-corner arithmetic of plane_box.  Only evaluated when the body's bounding
// sphere already reaches the plane: long pads (the SO-101 jaws) make that sphere test fire ten times more often than contact.
GG_HD bool pads_reach_plane(const GgModel& M, int b, const M3& Rb, V3 pb) {
    bool hit = false;
    for (int g = 0; g < M.n_pads; g++) {
        if (M.pad_body[g] != b) continue;
        const V3 c = pb + mul(Rb, ld3(M.pad_pos[g]));
        const M3 R = mmul(Rb, qmat(ldq(M.pad_quat[g])));
        const V3 h = ld3(M.pad_half[g]);
        const float ext = fabsf(R.c0.z) * h.x + fabsf(R.c1.z) * h.y + fabsf(R.c2.z) * h.z;
        hit = hit || (c.z - ext <= M.plane_z + 1e-5f);
    }
    return hit;
}

// Two oriented boxes: the six face axes of the separating-axis test (the nine edge-edge axes are skipped: a miss on any axis
// proves separation, so testing fewer axes only errs towards "may touch"); 1e-5 m margin against the narrow phase's own arithmetic.
GG_HD bool boxes_may_touch(V3 cA, const M3& RA, V3 hA, V3 cB, const M3& RB, V3 hB) {
    const V3 t = cB - cA;
    bool sep = false;
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        const V3 ua = col(RA, ax), ub = col(RB, ax);
        const float ra = comp(hA, ax) + fabsf(dot(ua, RB.c0)) * hB.x + fabsf(dot(ua, RB.c1)) * hB.y + fabsf(dot(ua, RB.c2)) * hB.z;
        sep = sep || (fabsf(dot(t, ua)) > ra + 1e-5f);
        const float rb = comp(hB, ax) + fabsf(dot(ub, RA.c0)) * hA.x + fabsf(dot(ub, RA.c1)) * hA.y + fabsf(dot(ub, RA.c2)) * hA.z;
        sep = sep || (fabsf(dot(t, ub)) > rb + 1e-5f);
    }
    return !sep;
}

// pads of body b against the cube, two conservative refinements of the sphere-sphere test:
//   (1) distance from the cube centre to the pad box (closest point of an oriented box) against the cube's bounding-sphere radius
//   (2) boxes_may_touch
GG_HD bool pads_reach_cube(const GgModel& M, int b, const M3& Rb, V3 pb, V3 cp, const M3& cR, V3 ch) {
    const float rc = sqrtf(dot(ch, ch)) + 1e-5f;
    bool hit = false;
    for (int g = 0; g < M.n_pads; g++) {
        if (M.pad_body[g] != b) continue;
        const V3 c = pb + mul(Rb, ld3(M.pad_pos[g]));
        const M3 R = mmul(Rb, qmat(ldq(M.pad_quat[g])));
        const V3 h = ld3(M.pad_half[g]);
        const V3 l = mulT(R, cp - c);
        const V3 e = mk(fmaxf(fabsf(l.x) - h.x, 0.f), fmaxf(fabsf(l.y) - h.y, 0.f), fmaxf(fabsf(l.z) - h.z, 0.f));
        if (dot(e, e) > rc * rc) continue;
        hit = hit || boxes_may_touch(c, R, h, cp, cR, ch);
    }
    return hit;
}

// conservative body-level broad phase: does any pad-carrying body's bounding sphere reach the plane or the cube?
template <int NA, int NF>
GG_HD bool pads_may_touch(const GgModel& M, const RobotDyn<NA, NF>& D, V3 cp, Q4 cq) {
    V3 ch = ld3(M.cube_half);
    const float rc = sqrtf(dot(ch, ch));
    bool any = false;
#pragma unroll
    for (int b = 0; b < 1 + NF; b++) {
        const float r = M.pad_bound[b][3];
        const M3& Rb = b == 0 ? D.lR : D.fR[b == 0 ? 0 : b - 1];
        V3 pb = b == 0 ? D.lp : D.fp[b == 0 ? 0 : b - 1];
        V3 c = pb + mul(Rb, ld3(M.pad_bound[b]));
        V3 d = c - cp;
        bool hit = false;
        if (r >= 0.f && dot(d, d) <= (r + rc) * (r + rc)) hit = pads_reach_cube(M, b, Rb, pb, cp, qmat(cq), ch);   // rare: refine the sphere tests
        if (r >= 0.f && !hit && c.z - r <= M.plane_z) hit = pads_reach_plane(M, b, Rb, pb);
        any = any || hit;
    }
    return any;
}

// ---------------------------------------------------------------------------------------------
// Building blocks of one substep (A.2): dynamics at t, collision, constraint solve, semi-implicit Euler.
// env_substep composes them for one env in one thread (host harness); the CUDA kernel runs them in three
// role warps per 32 envs (gg_kernels.cu).
// ---------------------------------------------------------------------------------------------
// cube block: diagonal inertia, free (unconstrained) acceleration and velocity vector [v_world, w_body]
GG_HD void cube_free_acc(const GgModel& M, V3 cv, V3 cw, float* Mc, float* as_c, float* qd_c) {
    Mc[0] = M.cube_mass; Mc[1] = M.cube_mass; Mc[2] = M.cube_mass;
    Mc[3] = M.cube_inertia[0]; Mc[4] = M.cube_inertia[1]; Mc[5] = M.cube_inertia[2];
    V3 Iw = mk(Mc[3] * cw.x, Mc[4] * cw.y, Mc[5] * cw.z);
    V3 gyro = cross(cw, Iw);
    as_c[0] = 0.f; as_c[1] = 0.f; as_c[2] = -M.gravity;
    as_c[3] = -gg_div(gyro.x, Mc[3]); as_c[4] = -gg_div(gyro.y, Mc[4]); as_c[5] = -gg_div(gyro.z, Mc[5]);
    qd_c[0] = cv.x; qd_c[1] = cv.y; qd_c[2] = cv.z; qd_c[3] = cw.x; qd_c[4] = cw.y; qd_c[5] = cw.z;
}

// joint-limit rows (at most one per dof); returns the number of violated limits
template <int NU>
GG_HD int joint_limit_rows(const GgModel& M, const float* q, const float* qd, float* lsg, float* laref, float* lD) {
    int nlim = 0;
#pragma unroll
    for (int d = 0; d < NU; d++) {
        float plo = q[d] - M.lim_lo[d], phi = M.lim_hi[d] - q[d];
        float sg = plo < 0.f ? 1.f : (phi < 0.f ? -1.f : 0.f);
        float pos = plo < 0.f ? plo : phi;
        lsg[d] = sg; laref[d] = 0.f; lD[d] = 0.f;
        // every consumer guards laref / lD with lsg != 0, so dofs that violate no limit in ANY env of the warp skip the
        // impedance evaluation (warp-uniform branch; typically only the gripper dofs sit on a limit)
        if (GG_WARP_ANY(sg != 0.f)) {
            float imp, aref;
            imp_aref(M, pos, sg * qd[d], imp, aref);
            laref[d] = aref;
            lD[d] = gg_rcp(fmaxf(M.minval, (1.f - imp) / imp * M.dof_invweight[d]));
        }
        nlim += sg != 0.f;
    }
    return nlim;
}

// cube corners touching the plane (first <= 4 in corner-index order, same points as plane_box)
GG_HD int cube_plane_points(const GgModel& M, V3 cp, const M3& cR, V3* pc, float* dc) {
    int ncp = 0;
    V3 half = ld3(M.cube_half);
#pragma unroll
    for (int c = 0; c < 4; c++) { pc[c] = cp; dc[c] = 0.f; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        V3 l = mk((i & 1) ? half.x : -half.x, (i & 2) ? half.y : -half.y, (i & 4) ? half.z : -half.z);
        V3 w = cp + mul(cR, l);
        float dist = w.z - M.plane_z;
        if (dist <= 0.f && ncp < 4) {
            V3 pp = mk(w.x, w.y, w.z - 0.5f * dist);
#pragma unroll
            for (int c = 0; c < 4; c++) if (c == ncp) { pc[c] = pp; dc[c] = -dist; }
            ncp++;
        }
    }
    return ncp;
}

// cube block on the decoupled path: plane contacts only
template <class CS>
GG_HD void cube_block_solve(const GgModel& M, V3 cp, const M3& cR, const float* Mc, const float* as_c, const float* qd_c,
                            const float* warm_c, float* a_c, int iterations, float tol, CS& cs, StepInfo& info) {
    V3 pc[4]; float dc[4];
    const int ncp = cube_plane_points(M, cp, cR, pc, dc);
    info.ncon = ncp; info.iters = 0; info.grad = 0.f;
    if (ncp == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) a_c[i] = as_c[i];
    } else {
        solve_cube_plane(M, Mc, as_c, qd_c, cp, cR, pc, dc, ncp, warm_c, a_c, iterations, tol, cs, info.iters, info.grad);
    }
}

// Joint-limit rows are unit rows (J = +-e_d), so for a GIVEN active set A the minimiser of the robot block has a closed form through
// the factor of the mass matrix the caller already has (Woodbury):
//     lambda = (D_A^-1 + E^T M^-1 E)^-1 (a_s - s r)_A,   a = a_s - M^-1 E lambda          (E = [e_d], d in A)
// and A is right iff row d is violated at a exactly for d in A (KKT of the convex problem: then a IS the minimiser).  With at most
// two active rows (the two fingers riding their stops is the common case) this costs |A| triangular solve pairs instead of a Newton
// iteration with its own 9 x 9 factorisation, mass-matrix products and line search.  The active set starts from the rows the
// unconstrained solution a_s violates (right at the first pass in every case of the random-target workload; the previous step's
// solution is a worse guess when the targets change every step) and is corrected from the solution; more than two active rows, or no
// fixed point after three passes, fall back to Newton.
template <int NU> GG_HD float sel_dyn(const float* v, int d) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < NU; i++) r = i == d ? v[i] : r;
    return r;
}
template <int NU>
GG_HD bool solve_robot_limits_direct(const float* L, const float* dinv, const float* as, const float* lsg, const float* laref, const float* lD,
                                     float* a, int& passes) {
    unsigned act = 0u;
#pragma unroll
    for (int d = 0; d < NU; d++) act |= (lsg[d] != 0.f && lsg[d] * as[d] - laref[d] < 0.f) ? 1u << d : 0u;
    for (passes = 1; passes <= 3; passes++) {
        if (popc32(act) > 2) return false;
        int d0 = -1, d1 = -1;
#pragma unroll
        for (int d = NU - 1; d >= 0; d--) if ((act >> d) & 1u) { d1 = d0; d0 = d; }      // d0 < d1: the active dofs
#pragma unroll
        for (int i = 0; i < NU; i++) a[i] = as[i];
        if (d0 >= 0) {
            float b[NU], z0[NU];
#pragma unroll
            for (int i = 0; i < NU; i++) b[i] = i == d0 ? 1.f : 0.f;
            chol_solve_packed<NU>(L, dinv, b, z0);
            const float c0 = sel_dyn<NU>(as, d0) - sel_dyn<NU>(lsg, d0) * sel_dyn<NU>(laref, d0);
            const float h00 = gg_rcp(sel_dyn<NU>(lD, d0)) + sel_dyn<NU>(z0, d0);
            if (d1 < 0) {
                const float lam0 = gg_div(c0, h00);
#pragma unroll
                for (int i = 0; i < NU; i++) a[i] -= z0[i] * lam0;
            } else {
                float z1[NU];
#pragma unroll
                for (int i = 0; i < NU; i++) b[i] = i == d1 ? 1.f : 0.f;
                chol_solve_packed<NU>(L, dinv, b, z1);
                const float c1 = sel_dyn<NU>(as, d1) - sel_dyn<NU>(lsg, d1) * sel_dyn<NU>(laref, d1);
                const float h11 = gg_rcp(sel_dyn<NU>(lD, d1)) + sel_dyn<NU>(z1, d1), h01 = sel_dyn<NU>(z0, d1);
                const float idet = gg_rcp(h00 * h11 - h01 * h01);
                const float lam0 = (c0 * h11 - c1 * h01) * idet, lam1 = (c1 * h00 - c0 * h01) * idet;
#pragma unroll
                for (int i = 0; i < NU; i++) a[i] -= z0[i] * lam0 + z1[i] * lam1;
            }
        }
        unsigned now = 0u;
#pragma unroll
        for (int d = 0; d < NU; d++) now |= (lsg[d] != 0.f && lsg[d] * a[d] - laref[d] < 0.f) ? 1u << d : 0u;
        if (now == act) return true;
        act = now;
    }
    return false;
}

// robot block on the decoupled path: joint-limit rows only (L, dinv: packed Cholesky factor of Mt)
template <int NU>
GG_HD void robot_block_solve(const float* Mt, const float* L, const float* dinv, const float* as, int nlim, const float* lsg, const float* laref,
                             const float* lD, const float* warm, float* a, int iterations, float tol, StepInfo& info) {
    info.ncon = 0; info.iters = 0; info.grad = 0.f;
    if (nlim == 0) {
#pragma unroll
        for (int i = 0; i < NU; i++) a[i] = as[i];
    } else {
#ifndef GG_NO_DIRECT_LIMITS
        int passes = 0;
        if (solve_robot_limits_direct<NU>(L, dinv, as, lsg, laref, lD, a, passes)) { info.iters = passes; return; }
#endif
        solve_robot_limits<NU>(Mt, as, lsg, laref, lD, warm, a, iterations, tol, 1.f, info.iters, info.grad);
    }
}

// coupled path: full narrow phase + Newton on all NU + 6 dofs
template <int NA, int NF, bool FH>
GG_HD void coupled_block_solve(const GgModel& M, const RobotDyn<NA, NF>& D, const EnvState<NA + NF>& s, const M3& cR, const float* Mc,
                               const float* as, const float* qdv, int nlim, const float* lsg, const float* laref, const float* lD,
                               float* a, int iterations, float tol, StepInfo& info) {
    constexpr int NV = NA + NF + 6;
    Contacts C;
    C.count = 0;
    const int maxc = M.max_contacts < GG_MAX_CONTACTS ? M.max_contacts : GG_MAX_CONTACTS;
    plane_box(M.plane_z, s.cp, cR, ld3(M.cube_half), 0, maxc, C);
    collide_pads<NA, NF>(M, D, s.cp, cR, C);
    info.ncon = C.count; info.iters = 0; info.grad = 0.f;
    if (nlim == 0 && C.count == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) a[i] = as[i];
    } else {
        solve_coupled<NA, NF, FH>(M, D, s, cR, Mc, as, qdv, lsg, laref, lD, C, a, iterations, tol, 1.f, info.iters, info.grad);
    }
}

template <int NU> GG_HD void integrate_robot(const GgModel& M, float* q, float* qd, float* qacc, const float* a) {
#pragma unroll
    for (int i = 0; i < NU; i++) { qacc[i] = a[i]; qd[i] += a[i] * M.dt; q[i] += qd[i] * M.dt; }
}

GG_HD void integrate_cube(const GgModel& M, V3& cp, Q4& cq, V3& cv, V3& cw, float* qacc_c, const float* a_c) {
    const float dt = M.dt;
#pragma unroll
    for (int i = 0; i < 6; i++) qacc_c[i] = a_c[i];
    cv = cv + mk(a_c[0], a_c[1], a_c[2]) * dt;
    cw = cw + mk(a_c[3], a_c[4], a_c[5]) * dt;
    cp = cp + cv * dt;
    V3 w = cw * dt;
    float ang = sqrtf(dot(w, w));
    Q4 dq = mkq(1.f, 0.f, 0.f, 0.f);
    if (ang > 0.f) {
        float sn, cs;
        gg_sincos(0.5f * ang, &sn, &cs);
        float k = gg_div(sn, ang);
        dq = mkq(cs, w.x * k, w.y * k, w.z * k);
    }
    Q4 nq = qmul(cq, dq);
    float inv = gg_rsqrt(nq.w * nq.w + nq.x * nq.x + nq.y * nq.y + nq.z * nq.z);
    cq = mkq(nq.w * inv, nq.x * inv, nq.y * inv, nq.z * inv);
}

template <int NA, int NF, bool FH>
GG_HD void env_substep(const GgModel& M, const GgChain& C, EnvState<NA + NF>& s, const float* target, int iterations, float tol, StepInfo& info) {
    constexpr int NU = NA + NF, NV = NU + 6;
    RobotDyn<NA, NF> D;
    robot_dynamics<NA, NF, FH>(M, C, s.q, s.qd, target, D);
    const M3 cR = qmat(s.cq);

    // ---- smooth (unconstrained) acceleration a_s
    float Mc[6], as[NV], qdv[NV];
    float L[NU * (NU + 1) / 2], dinv[NU];
    {
#pragma unroll
        for (int i = 0; i < NU * (NU + 1) / 2; i++) L[i] = D.Mt[i];
        chol_packed<NU>(L, dinv);
        chol_solve_packed<NU>(L, dinv, D.fs, as);
#pragma unroll
        for (int i = 0; i < NU; i++) qdv[i] = s.qd[i];
        cube_free_acc(M, s.cv, s.cw, Mc, as + NU, qdv + NU);
    }
    float lsg[NU], laref[NU], lD[NU];
    const int nlim = joint_limit_rows<NU>(M, s.q, s.qd, lsg, laref, lD);

    float a[NV];
    if (!pads_may_touch<NA, NF>(M, D, s.cp, s.cq)) {
        // ---- decoupled fast path: robot block (limits only) and cube block (plane contacts only)
        StepInfo ir, ic;
        CubeConLocal cs;
        robot_block_solve<NU>(D.Mt, L, dinv, as, nlim, lsg, laref, lD, s.qacc, a, iterations, tol, ir);
        cube_block_solve(M, s.cp, cR, Mc, as + NU, qdv + NU, s.qacc + NU, a + NU, iterations, tol, cs, ic);
        info.ncon = ic.ncon;
        info.iters = ic.iters > ir.iters ? ic.iters : ir.iters;
        info.grad = fmaxf(ir.grad, ic.grad);
    } else {
        coupled_block_solve<NA, NF, FH>(M, D, s, cR, Mc, as, qdv, nlim, lsg, laref, lD, a, iterations, tol, info);
    }

    // ---- semi-implicit Euler
    integrate_robot<NU>(M, s.q, s.qd, s.qacc, a);
    integrate_cube(M, s.cp, s.cq, s.cv, s.cw, s.qacc + NU, a + NU);
}

// ---------------------------------------------------------------------------------------------
// Observation / reward / success of the reference task classes
// (franka/cube_pick.py:130-157; env.py:63-65).  agent: 7 + NF floats, envs: 11 floats.
// ---------------------------------------------------------------------------------------------
template <int NA, int NF>
GG_HD void env_observe(const GgModel& M, const EnvState<NA + NF>& s, float* agent, float* envs, float& reward, bool& success) {
    V3 ep; Q4 eq;
    {
        constexpr int KS = (NA - 1) / 2;
        Q4 ql, tq; V3 pl, tp; M3 Rl;
        robot_chain_head<NA>(M, s.q, KS, ql, pl, Rl);
        robot_chain_tail<NA>(M, s.q + KS, KS, tp, tq);
        compose_eef(ql, pl, Rl, tp, tq, ep, eq);
    }
    agent[0] = ep.x; agent[1] = ep.y; agent[2] = ep.z;
    agent[3] = eq.w; agent[4] = eq.x; agent[5] = eq.y; agent[6] = eq.z;
#pragma unroll
    for (int f = 0; f < NF; f++) agent[7 + f] = s.q[NA + f];
    envs[0] = s.cp.x; envs[1] = s.cp.y; envs[2] = s.cp.z;
    envs[3] = s.cq.w; envs[4] = s.cq.x; envs[5] = s.cq.y; envs[6] = s.cq.z;
    V3 d = ep - s.cp;
    envs[7] = d.x; envs[8] = d.y; envs[9] = d.z;
    envs[10] = sqrtf(dot(d, d));
    success = s.cp.z > M.success_z;
    reward = success ? 1.f : 0.f;
}

// reset one env: home pose, zero velocities / warm start, cube at (x, y, spawn_z) with the spawn quaternion
template <int NU>
GG_HD void env_reset(const GgModel& M, EnvState<NU>& s, float x, float y) {
#pragma unroll
    for (int i = 0; i < NU; i++) { s.q[i] = M.home_q[i]; s.qd[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < NU + 6; i++) s.qacc[i] = 0.f;
    s.cp = mk(x, y, M.spawn_z);
    s.cq = ldq(M.spawn_quat);
    s.cv = mk(0.f, 0.f, 0.f); s.cw = mk(0.f, 0.f, 0.f);
}

// device-side spawn draw: Philox counter (global env id, episode, 0, 0), key = seed
GG_HD void spawn_draw(const GgModel& M, uint64_t seed, uint32_t gid, uint32_t episode, float& x, float& y) {
    uint32_t o[4];
    philox4x32_10(gid, episode, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    x = fmaf(u01(o[0]), M.spawn_hi[0] - M.spawn_lo[0], M.spawn_lo[0]);
    y = fmaf(u01(o[1]), M.spawn_hi[1] - M.spawn_lo[1], M.spawn_lo[1]);
}

}  // namespace gg
