// Multi-object scenes: FrankaCubeStackEmptyBatch, gym_genesis/tasks/franka/cube_stack.py:27-240 --
// the Panda, cube_1, cube_2 and three distractor cubes on the ground plane, cube-cube box-box contacts enabled.
//
// Same algorithm as the CubePick path (gg_env.cuh / gg_coop.cuh): the unique minimiser of
//     1/2 (a - a_s)^T Mt (a - a_s) + sum_active 1/2 D (J a - aref)^2        
// by Newton with exact line search, on NU + 6 K dofs.  Three stages per env.step, each with its own thread mapping:
//   stack_prep    thread per env   robot dynamics (reuses robot_dynamics / chol_packed / joint_limit_rows of gg_env.cuh), pad
//                                  broad phase against every cube and the plane; when no pad can touch anything the robot block
//                                  decouples and is solved here in registers (robot_block_solve).  Publishes a StackWork record.
//   stack_solve   WARP per env     narrow phase (lane = geom pair: cube-plane, pad-plane, pad-cube, cube-cube; ordered compaction),
//                                  per-contact rows (lane = contact, 3 per lane at the 96-contact cap), pair stiffness in
//                                  spatial-force space, projection onto the dofs of each body pair, dense Cholesky of the
//                                  coupled block in shared memory, line search with butterfly reductions.
//   stack_finish  thread per env   semi-implicit Euler of robot and cubes, observations, reward / flags.
// The warp-level code is written against a small warp-context interface (sync / shfl / sum / max / any / ballot / scan) so that the
// CPU-tier test harness (tests/hostsim) can run the SAME source with 32 host threads per emulated warp.
//
// Body numbering inside an env: 0 world (plane), 1 last arm link (hand), 2 + f finger f, 4 + k cube k.
// Contact pair (A, B), A < B, normal from A to B; pair id: plane-X, then robot body-cube, then cube-cube (row-major).
#pragma once
#include "gg_env.cuh"

#if defined(__CUDACC__)
#define GG_WD __device__ __forceinline__
#else
#define GG_WD inline
#endif

namespace gg {

constexpr int SK_CAP = GG_STACK_MAX_CONTACTS;
constexpr int SK_CPL = (SK_CAP + 31) / 32;     // contacts per lane
constexpr int SK_FS = 19;                      // row stride of the per-contact force table (odd: conflict free for lane = contact)

template <int K> struct ObjState { V3 p[K]; Q4 q[K]; V3 v[K], w[K]; };

// published by stack_prep, consumed by stack_solve (plain floats: moved through global memory with coalesced copies)
template <int NA, int NF> struct StackWork {
    static constexpr int NU = NA + NF, NFX = NF > 0 ? NF : 1;
    float Mt[NU * (NU + 1) / 2];
    float as[NU], qd[NU], warm[NU], lsg[NU], laref[NU], lD[NU], arob[NU];
    float jk[NA][3], jo[NA][3];
    float fk[NFX][3], fo[NFX][3];
    float lR[9], lp[3];
    float fR[NFX][9], fp[NFX][3];
    float amask;         // bodies of the cooperative solve: bit 0 robot (a pad may touch a cube or the plane), bit 1 + k cube k
                         // (within reach of a pad or of another cube); everything else was solved in registers by stack_prep
    float active;        // 1: this env runs a substep
    float iso_iters;     // largest Newton iteration count of the register solves
    float iso_ncon;      // plane contacts of the cubes solved in registers
};

template <int NA, int NF, int K> struct StackIn {
    StackWork<NA, NF> W;
    float obj[K][13];     // pos3 quat4 linvel3 angvel3 (body frame)
    float warm_obj[K][6];
    float aobj[K][6];     // acceleration of the cubes solved in registers (isolated: 6 x 6 block with <= 4 plane contacts)
};

template <int NU, int K> struct StackBuf {
    static constexpr int NV = NU + 6 * K, NH = NV * (NV + 1) / 2, NB = 4 + K, NPAIR = (NB - 1) + 3 * K + K * (K - 1) / 2;
    float F[SK_CAP][SK_FS];   // spatial forces (moment about O, direction) of n, t1, t2; first p, n while colliding
    float Wc[8][SK_CAP];      // per-iteration weights Wnn Wn1 W11 Wn2 W22 gn g1 g2, component-major
    float depth[SK_CAP];
    int code[SK_CAP];         // pair id | A << 8 | B << 16
    float TB[NV][6];          // twist of every dof about O (angular, linear)
    float T[3][NB][6];        // body twists of qd / warm start / a_smooth (slot 2 reused for the search direction)
    float Phi[NPAIR][21], gF[NPAIR][6];
    int pairAB[NPAIR];        // -1: pair has no contact
    float Y[NV][6];
    float H[NH];
    float dinv[NV], vx[NV];
    int lo[NV];               // dof -> local (compacted) row index of the active dofs
    int plist[NPAIR];         // ids of the body pairs in contact
    int pstart[NPAIR], pcnt[NPAIR];   // range of every pair in perm
    int perm[SK_CAP];         // contact indices grouped by body pair
    float part[5 * 42];       // partial pair-stiffness sums of the five contact subsets
    int tlist[SK_CAP];        // geom pairs that survive the narrow phase's cull, in task order (<= 5 + 12 + 60 + 10 tasks)
};

// ---- body / pair bookkeeping ------------------------------------------------------------------------------------------
template <int NA, int NF> GG_HD bool sk_in(int b, int d) {
    constexpr int NU = NA + NF;
    if (b == 0) return false;
    if (b < 4) return d < NA || (b >= 2 && b - 2 < NF && d == NA + b - 2);
    const int r = d - NU - 6 * (b - 4);
    return r >= 0 && r < 6;
}
template <int NA, int NF> GG_HD float sk_sign(int A, int B, int d) {
    return (sk_in<NA, NF>(B, d) ? 1.f : 0.f) - (sk_in<NA, NF>(A, d) ? 1.f : 0.f);
}
template <int NA, int NF> GG_HD int sk_ndof(int b) { return b == 0 ? 0 : (b == 1 ? NA : (b < 4 ? NA + 1 : 6)); }
template <int NA, int NF> GG_HD int sk_dof(int b, int i) { return b < 4 ? (i < NA ? i : NA + b - 2) : NA + NF + 6 * (b - 4) + i; }
template <int K> GG_HD int sk_pair_id(int A, int B) {
    constexpr int NB = 4 + K;
    if (A == 0) return B - 1;
    if (A < 4) return (NB - 1) + (A - 1) * K + (B - 4);
    const int i = A - 4, j = B - 4;
    return (NB - 1) + 3 * K + i * K - i * (i + 1) / 2 + (j - i - 1);
}
GG_HD int sk_code(int pid, int A, int B) { return pid | (A << 8) | (B << 16); }

// row of packed entry e (row-major lower triangle)
GG_HD int sk_tri_row(int e) {
    int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
    while (i * (i + 1) / 2 > e) i--;
    while ((i + 1) * (i + 2) / 2 <= e) i++;
    return i;
}

GG_HD float sk_dot6(const float* a, const float* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 1: thread per env
// ---------------------------------------------------------------------------------------------------------------------
// conservative body-level broad phase: does any pad-carrying body's bounding sphere reach the plane / cube k?
template <int NA, int NF, int K>
GG_HD void stack_pad_reach(const GgModel& M, const RobotDyn<NA, NF>& D, const V3* cp, const Q4* cq, bool& plane, unsigned& cubes) {
    const V3 ch = ld3(M.cube_half);
    const float rc = sqrtf(dot(ch, ch));
    plane = false; cubes = 0u;
#pragma unroll
    for (int b = 0; b < 1 + NF; b++) {
        const float r = M.pad_bound[b][3];
        if (r < 0.f) continue;
        const M3& Rb = b == 0 ? D.lR : D.fR[b == 0 ? 0 : b - 1];
        const V3 pb = b == 0 ? D.lp : D.fp[b == 0 ? 0 : b - 1];
        const V3 c = pb + mul(Rb, ld3(M.pad_bound[b]));
        if (!plane && c.z - r <= M.plane_z) plane = pads_reach_plane(M, b, Rb, pb);     // refine the sphere test (gg_env.cuh)
#pragma unroll
        for (int k = 0; k < K; k++) {
            const V3 d = c - cp[k];
            if (!((cubes >> k) & 1u) && dot(d, d) <= (r + rc) * (r + rc) && pads_reach_cube(M, b, Rb, pb, cp[k], qmat(cq[k]), ch)) cubes |= 1u << k;
        }
    }
}

// The prep stage is independent pieces (the CUDA kernel runs them in three role warps of the same block, lane = env):
//   stack_prep_robot : CRBA, Cholesky, limit rows, pad reach, robot-block solve when no pad can touch anything -> In.W (all but the
//                      amask / iso_* words), returns the robot's amask bits (bit 0 robot, bit 1 + k cube k within reach of a pad);
//                      the bias / PD half of the dynamics pass (robot_rnea) arrives through the FSP functor
//   stack_prep_cubes : cube-cube reach, register solves of the isolated cubes -> In.aobj, returns the cube bits and statistics
// The pad reach is needed by both; the cube half recomputes it from the pad-body frames (robot_pad_frames, one chain walk).
// FSP: provider of fs = clamp(PD) - damping qd - bias.  The host harness computes it in place (robot_rnea); the CUDA kernel's ROBOT
// warp receives it from the RNEA warp, which walks the same chain concurrently (the two halves of the dynamics pass are independent).
template <int NA, int NF, bool FH> struct FsInline {
    const GgModel* M; const GgChain* C; const float* q; const float* qd; const float* target;
    GG_HD void operator()(float* fs) const { robot_rnea<NA, NF, FH>(*M, *C, q, qd, target, fs); }
};

// RMP: provider of the robot's amask bits (bit 0: a pad may touch the plane or a cube, bit 1 + k: cube k within reach of a pad).
// The host harness evaluates the pad reach in place from the frames of this pass; the CUBES warp of the CUDA kernel evaluates it
// once for both roles (one source of truth: the two roles can never disagree about who is coupled) and hands it over.
template <int NA, int NF, int K> struct RmaskInline {
    const GgModel* M; const V3* cube_pos; const Q4* cube_quat;
    GG_HD unsigned operator()(const RobotDyn<NA, NF>& D) const {
        bool pad_plane; unsigned pad_cubes;
        stack_pad_reach<NA, NF, K>(*M, D, cube_pos, cube_quat, pad_plane, pad_cubes);
        return ((pad_plane || pad_cubes != 0u) ? 1u : 0u) | (pad_cubes << 1);
    }
};

template <int NA, int NF, bool FH, int K, class FSP, class RMP>
GG_HD unsigned stack_prep_robot(const GgModel& M, const GgChain& C, const float* q, const float* qd, const float* warm, int iterations, float tol, bool active,
                                StackWork<NA, NF>& W, int& rob_iters, FSP& fsp, RMP& rmp) {
    constexpr int NU = NA + NF;
    RobotDyn<NA, NF> D;
    {
        LinkStoreLocal<NA> ls;
        robot_crba<NA, NF, FH>(M, C, q, D, ls);
    }
    float L[NU * (NU + 1) / 2], dinv[NU];
#pragma unroll
    for (int i = 0; i < NU * (NU + 1) / 2; i++) L[i] = D.Mt[i];
    chol_packed<NU>(L, dinv);
    float lsg[NU], laref[NU], lD[NU], a[NU];
    const int nlim = joint_limit_rows<NU>(M, q, qd, lsg, laref, lD);
    float as[NU];
    fsp(D.fs);
    const unsigned rmask = rmp(D);
    chol_solve_packed<NU>(L, dinv, D.fs, as);
    StepInfo ir; ir.ncon = 0; ir.iters = 0; ir.grad = 0.f;
#pragma unroll
    for (int i = 0; i < NU; i++) a[i] = as[i];
    if (active && !(rmask & 1u)) robot_block_solve<NU>(D.Mt, L, dinv, as, nlim, lsg, laref, lD, warm, a, iterations, tol, ir);
    rob_iters = ir.iters;
#pragma unroll
    for (int i = 0; i < NU * (NU + 1) / 2; i++) W.Mt[i] = D.Mt[i];
#pragma unroll
    for (int i = 0; i < NU; i++) {
        W.as[i] = as[i]; W.qd[i] = qd[i]; W.warm[i] = warm[i]; W.lsg[i] = lsg[i]; W.laref[i] = laref[i]; W.lD[i] = lD[i]; W.arob[i] = a[i];
    }
#pragma unroll
    for (int i = 0; i < NA; i++) {
        W.jk[i][0] = D.jk[i].x; W.jk[i][1] = D.jk[i].y; W.jk[i][2] = D.jk[i].z;
        W.jo[i][0] = D.jo[i].x; W.jo[i][1] = D.jo[i].y; W.jo[i][2] = D.jo[i].z;
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
        W.fk[f][0] = D.fk[f].x; W.fk[f][1] = D.fk[f].y; W.fk[f][2] = D.fk[f].z;
        W.fo[f][0] = D.fo[f].x; W.fo[f][1] = D.fo[f].y; W.fo[f][2] = D.fo[f].z;
        const M3& R = D.fR[f];
        W.fR[f][0] = R.c0.x; W.fR[f][1] = R.c0.y; W.fR[f][2] = R.c0.z; W.fR[f][3] = R.c1.x; W.fR[f][4] = R.c1.y; W.fR[f][5] = R.c1.z;
        W.fR[f][6] = R.c2.x; W.fR[f][7] = R.c2.y; W.fR[f][8] = R.c2.z;
        W.fp[f][0] = D.fp[f].x; W.fp[f][1] = D.fp[f].y; W.fp[f][2] = D.fp[f].z;
    }
    W.lR[0] = D.lR.c0.x; W.lR[1] = D.lR.c0.y; W.lR[2] = D.lR.c0.z; W.lR[3] = D.lR.c1.x; W.lR[4] = D.lR.c1.y; W.lR[5] = D.lR.c1.z;
    W.lR[6] = D.lR.c2.x; W.lR[7] = D.lR.c2.y; W.lR[8] = D.lR.c2.z;
    W.lp[0] = D.lp.x; W.lp[1] = D.lp.y; W.lp[2] = D.lp.z;
    W.active = active ? 1.f : 0.f;
    return rmask;
}

// pad_cubes: bit k = cube k within reach of a pad.  Returns the cube bits of amask (bit 1 + k): cubes within reach of a pad or of
// another cube.
template <int K>
GG_HD unsigned stack_cube_reach(const GgModel& M, const ObjState<K>& O, unsigned pad_cubes) {
    unsigned cmask = pad_cubes << 1;
    const V3 ch = ld3(M.cube_half);
    const float r2 = 4.f * dot(ch, ch);
#pragma unroll
    for (int i = 0; i < K; i++)
#pragma unroll
        for (int j = i + 1; j < K; j++) {
            const V3 d = O.p[i] - O.p[j];
            // bounding spheres first; where they fire (resting neighbours sit one cube width apart, well inside 2 r), the six
            // face axes of the SAT decide: cubes that were pushed apart and now rest a hair from each other stay decoupled
            if (dot(d, d) <= r2 && boxes_may_touch(O.p[i], qmat(O.q[i]), ch, O.p[j], qmat(O.q[j]), ch)) cmask |= (2u << i) | (2u << j);
        }
    return cmask;
}

// register solve of ONE isolated cube (6 x 6 block, <= 4 plane contacts): a_c, contacts, Newton iterations
GG_HD void stack_cube_isolated(const GgModel& M, V3 p, Q4 q, V3 v, V3 w, const float* warm_c, int iterations, float tol, bool solve,
                               float* a_c, int& ncon, int& iters) {
    float Mc[6], as_c[6], qd_c[6], w_c[6];
    cube_free_acc(M, v, w, Mc, as_c, qd_c);
#pragma unroll
    for (int i = 0; i < 6; i++) { a_c[i] = as_c[i]; w_c[i] = warm_c[i]; }
    ncon = 0; iters = 0;
    if (solve) {
        StepInfo ic;
        CubeConLocal cs;
        cube_block_solve(M, p, qmat(q), Mc, as_c, qd_c, w_c, a_c, iterations, tol, cs, ic);
        ncon = ic.ncon; iters = ic.iters;
    }
}

template <int K>
GG_HD unsigned stack_prep_cubes(const GgModel& M, const ObjState<K>& O, const float (*warm_obj)[6], unsigned pad_cubes, int iterations,
                                float tol, bool active, float (*aobj)[6], int& iso_iters, int& iso_ncon) {
    const unsigned cmask = stack_cube_reach<K>(M, O, pad_cubes);
    iso_iters = 0; iso_ncon = 0;
    for (int k = 0; k < K; k++) {
        int nc, it;
        stack_cube_isolated(M, O.p[k], O.q[k], O.v[k], O.w[k], warm_obj[k], iterations, tol, active && !((cmask >> (1 + k)) & 1u), aobj[k], nc, it);
        iso_ncon += nc;
        iso_iters = it > iso_iters ? it : iso_iters;
    }
    return cmask;
}

// both halves in one thread (host harness)
template <int NA, int NF, bool FH, int K>
GG_HD void stack_prep(const GgModel& M, const GgChain& C, const float* q, const float* qd, const float* warm, const float* target, const ObjState<K>& O,
                      int iterations, float tol, bool active, StackIn<NA, NF, K>& In) {
    int rob_iters, iso_iters, iso_ncon;
    FsInline<NA, NF, FH> fsp{&M, &C, q, qd, target};
    RmaskInline<NA, NF, K> rmp{&M, O.p, O.q};
    const unsigned rmask = stack_prep_robot<NA, NF, FH, K>(M, C, q, qd, warm, iterations, tol, active, In.W, rob_iters, fsp, rmp);
    const unsigned cmask = stack_prep_cubes<K>(M, O, In.warm_obj, rmask >> 1, iterations, tol, active, In.aobj, iso_iters, iso_ncon);
    In.W.amask = (float)(rmask | cmask);
    In.W.iso_iters = (float)(rob_iters > iso_iters ? rob_iters : iso_iters);
    In.W.iso_ncon = (float)iso_ncon;
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 2: warp per env
// ---------------------------------------------------------------------------------------------------------------------
GG_HD M3 sk_mat9(const float* r) {
    M3 R; R.c0 = mk(r[0], r[1], r[2]); R.c1 = mk(r[3], r[4], r[5]); R.c2 = mk(r[6], r[7], r[8]);
    return R;
}

// narrow phase; contacts come out in the serial order (cubes-plane, pads-plane, pad-major pad-cube, row-major cube-cube) --
// oracle.cpp: collide.  Three steps so that the expensive routine runs on full lanes only:
//   (1) every geom pair ("task") is decoded and CULLED by one lane: bounding spheres, then the six face axes of the SAT
//       (boxes_may_touch) / the exact box-plane reach -- cheap, ~100 instructions;
//   (2) the surviving task ids are compacted in task order into B.tlist;
//   (3) lane i re-decodes survivor i and runs plane_box / box_box: one round for up to 32 survivors (a grasp has ~12) instead of
//       one round per 32 TASKS (a grasp has 75), each of which paid the full 2 k-instruction box_box for its few live lanes.
// plane_box and box_box have ONE call site each (they are ~3 k instructions inlined; the solver is instruction-supply bound).
template <int NA, int NF, int K>
struct SkPair { int kind, code; V3 cA, hA, cB; M3 RA, RB; };      // kind: 0 none, 1 plane-box, 2 box-box

template <int NA, int NF, int K>
GG_HD SkPair<NA, NF, K> sk_decode(const GgModel& M, const StackIn<NA, NF, K>& In, unsigned amask, int np, int t) {
    SkPair<NA, NF, K> P;
    const V3 ch = ld3(M.cube_half);
    const float rc = sqrtf(dot(ch, ch));
    P.kind = 0; P.code = 0; P.cA = mk(0.f, 0.f, 0.f); P.hA = ch; P.cB = P.cA; P.RA = sk_mat9(In.W.lR); P.RB = P.RA;
    if (t < K) {
        const float* o = In.obj[t];
        if ((amask >> (1 + t)) & 1u) {
            P.kind = 1; P.cA = mk(o[0], o[1], o[2]); P.RA = qmat(mkq(o[3], o[4], o[5], o[6]));
            P.code = sk_code(sk_pair_id<K>(0, 4 + t), 0, 4 + t);
        }
    } else if (t < K + np + np * K) {
        t -= K;
        const bool plane = t < np;
        const int g = plane ? t : (t - np) / K;
        const int k = plane ? 0 : (t - np) % K;
        const int b = M.pad_body[g];
        const int fi = b - 1 < 0 ? 0 : (b - 1 < NF ? b - 1 : 0);
        const M3 Rb = b == 0 ? sk_mat9(In.W.lR) : sk_mat9(In.W.fR[fi]);
        const V3 pb = b == 0 ? ld3(In.W.lp) : ld3(In.W.fp[fi]);
        P.cA = pb + mul(Rb, ld3(M.pad_pos[g]));
        P.hA = ld3(M.pad_half[g]);
        const float rp = sqrtf(dot(P.hA, P.hA));
        if (plane) {
            if (P.cA.z - rp <= M.plane_z) {
                P.RA = mmul(Rb, qmat(ldq(M.pad_quat[g])));
                const float ext = fabsf(P.RA.c0.z) * P.hA.x + fabsf(P.RA.c1.z) * P.hA.y + fabsf(P.RA.c2.z) * P.hA.z;
                if (P.cA.z - ext <= M.plane_z + 1e-5f) { P.kind = 1; P.code = sk_code(sk_pair_id<K>(0, 1 + b), 0, 1 + b); }
            }
        } else {
            const float* o = In.obj[k];
            P.cB = mk(o[0], o[1], o[2]);
            const V3 dd = P.cA - P.cB;
            if (((amask >> (1 + k)) & 1u) && dot(dd, dd) <= (rp + rc) * (rp + rc)) {
                P.RA = mmul(Rb, qmat(ldq(M.pad_quat[g]))); P.RB = qmat(mkq(o[3], o[4], o[5], o[6]));
                if (boxes_may_touch(P.cA, P.RA, P.hA, P.cB, P.RB, ch)) { P.kind = 2; P.code = sk_code(sk_pair_id<K>(1 + b, 4 + k), 1 + b, 4 + k); }
            }
        }
    } else {
        t -= K + np + np * K;
        int i = 0;
        while (t >= K - 1 - i) { t -= K - 1 - i; i++; }
        const int j = i + 1 + t;
        const float* oi = In.obj[i];
        const float* oj = In.obj[j];
        P.cA = mk(oi[0], oi[1], oi[2]); P.cB = mk(oj[0], oj[1], oj[2]);
        const V3 dd = P.cA - P.cB;
        if (((amask >> (1 + i)) & (amask >> (1 + j)) & 1u) && dot(dd, dd) <= 4.f * rc * rc) {
            P.RA = qmat(mkq(oi[3], oi[4], oi[5], oi[6])); P.RB = qmat(mkq(oj[3], oj[4], oj[5], oj[6]));
            if (boxes_may_touch(P.cA, P.RA, ch, P.cB, P.RB, ch)) { P.kind = 2; P.code = sk_code(sk_pair_id<K>(4 + i, 4 + j), 4 + i, 4 + j); }
        }
    }
    return P;
}

template <int NA, int NF, int K, class WC>
GG_WD int stack_collide(const GgModel& M, int max_contacts, const StackIn<NA, NF, K>& In, unsigned amask, StackBuf<NA + NF, K>& B, WC& wc) {
    const int lane = wc.lane;
    const int maxc = max_contacts < SK_CAP ? max_contacts : SK_CAP;
    const V3 ch = ld3(M.cube_half);
    const int np = (amask & 1u) ? M.n_pads : 0;             // bodies outside the cooperative solve were handled by stack_prep
    const int nT = K + np + np * K + K * (K - 1) / 2;
    // (1) + (2): cull every task, compact the survivors in task order
    int nsurv = 0;
    for (int t0 = 0; t0 < nT; t0 += 32) {
        const int t = t0 + lane;
        int keep = 0;
        if (t < nT) keep = sk_decode<NA, NF, K>(M, In, amask, np, t).kind != 0 ? 1 : 0;
        const int incl = wc.scan_incl(keep);
        if (keep) B.tlist[nsurv + incl - 1] = t;
        nsurv += wc.shfli(incl, 31);
    }
    wc.sync();
    // (3): the narrow-phase routines on full lanes
    int total = 0;
    for (int s0 = 0; s0 < nsurv; s0 += 32) {
        Contacts8 C;
        C.count = 0;
        SkPair<NA, NF, K> P;
        P.kind = 0;
        if (s0 + lane < nsurv) P = sk_decode<NA, NF, K>(M, In, amask, np, B.tlist[s0 + lane]);
        if (wc.any(P.kind == 1)) { if (P.kind == 1) plane_box(M.plane_z, P.cA, P.RA, P.hA, P.code, 8, C); }
        if (wc.any(P.kind == 2)) { if (P.kind == 2) box_box(P.cA, P.RA, P.hA, P.cB, P.RB, ch, P.code, 8, C); }
        const int incl = wc.scan_incl(C.count);
        const int base = total + incl - C.count;
        for (int k = 0; k < C.count; k++) {
            const int c = base + k;
            if (c < maxc) {
                float* f = B.F[c];
                f[0] = C.p[k].x; f[1] = C.p[k].y; f[2] = C.p[k].z; f[3] = C.n[k].x; f[4] = C.n[k].y; f[5] = C.n[k].z;
                B.depth[c] = C.depth[k]; B.code[c] = C.code[k];
            }
        }
        total += wc.shfli(incl, 31);
        if (total > maxc) total = maxc;
    }
    wc.sync();
    return total;
}

// the solve over the ACTIVE bodies (W.amask != 0).  a_out[d * a_stride] receives the acceleration of the active dofs; the others
// keep the result of stack_prep's register solves (W.arob, In.aobj), which the caller has already stored there.
// Ownership: lane l (and l + 32 when more than 32 dofs are active -- a warp-uniform rarity) owns the l-th ACTIVE dof, i.e. row l of
// the packed Hessian; dd[] (registers) and B.lo translate between that compacted index and the dof number.
template <int NA, int NF, bool FH, int K, class WC>
GG_WD void stack_solve(const GgModel& M, int max_contacts, const StackIn<NA, NF, K>& In, StackBuf<NA + NF, K>& B, int iterations, float tol,
                       WC& wc, float* a_out, int a_stride, int& iters_out, int& ncon_out) {
    constexpr int NU = NA + NF, NV = NU + 6 * K, NB = 4 + K, ND = (NV + 31) / 32;
    typedef StackBuf<NU, K> BufT;
    const StackWork<NA, NF>& W = In.W;
    const int lane = wc.lane;
    const float mu = M.friction;
    const unsigned amask = (unsigned)W.amask;
    const bool cpl = (amask & 1u) != 0u;
    const int nrob = cpl ? NU : 0;
    int n = nrob;
#pragma unroll
    for (int k = 0; k < K; k++) n += ((amask >> (1 + k)) & 1u) ? 6 : 0;
    const bool wide = n > 32;                          // slot 1 of the per-lane arrays is in use
    // reference point of the spatial vectors: the first active cube (contacts of interest are near it), else cube 0
    int kref = 0;
#pragma unroll
    for (int k = K - 1; k >= 0; k--) if ((amask >> (1 + k)) & 1u) kref = k;
    const V3 O = mk(In.obj[kref][0], In.obj[kref][1], In.obj[kref][2]);

    // ---- contacts
    const int ncon = stack_collide<NA, NF, K>(M, max_contacts, In, amask, B, wc);
    ncon_out = ncon + (int)W.iso_ncon;
    for (int i = lane; i < BufT::NPAIR; i += 32) B.pairAB[i] = -1;
    wc.sync();
#pragma unroll
    for (int s = 0; s < SK_CPL; s++) {
        const int c = lane + 32 * s;
        if (c < ncon) {
            float* f = B.F[c];
            const V3 p = mk(f[0], f[1], f[2]) - O, nn = mk(f[3], f[4], f[5]);
            V3 t1, t2;
            tangent_basis(nn, t1, t2);
            const V3 pn = cross(p, nn), p1 = cross(p, t1), p2 = cross(p, t2);
            f[0] = pn.x; f[1] = pn.y; f[2] = pn.z; f[3] = nn.x; f[4] = nn.y; f[5] = nn.z;
            f[6] = p1.x; f[7] = p1.y; f[8] = p1.z; f[9] = t1.x; f[10] = t1.y; f[11] = t1.z;
            f[12] = p2.x; f[13] = p2.y; f[14] = p2.z; f[15] = t2.x; f[16] = t2.y; f[17] = t2.z;
            const int code = B.code[c];
            B.pairAB[code & 0xff] = code >> 8;         // same value from every contact of the pair
        }
    }

    // ---- per-dof data: lane owns active dofs number lane, lane + 32
    float Bd[ND][6], as_d[ND], qd_d[ND], warm_d[ND], Mc_d[ND];
    bool val[ND];
    int dd[ND];                                        // dof number of the owned row (0 when none)
    float lsg_d = 0.f, laref_d = 0.f, lD_d = 0.f;
    float Mrow[NU];
#pragma unroll
    for (int j = 0; j < NU; j++) Mrow[j] = 0.f;
#pragma unroll
    for (int s = 0; s < ND; s++) {
        const int l = lane + 32 * s;
#pragma unroll
        for (int k = 0; k < 6; k++) Bd[s][k] = 0.f;
        as_d[s] = 0.f; qd_d[s] = 0.f; warm_d[s] = 0.f; Mc_d[s] = 0.f;
        val[s] = l < n; dd[s] = 0;
        if (s > 0 && !wide) continue;
        if (!val[s]) continue;
        int d;
        if (l < nrob) d = l;
        else {
            const int r = (l - nrob) / 6, j = (l - nrob) - 6 * r;
            int kb = 0, seen = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const bool on = ((amask >> (1 + k)) & 1u) != 0u;
                if (on && seen == r) kb = k;
                seen += on ? 1 : 0;
            }
            d = NU + 6 * kb + j;
        }
        dd[s] = d;
        B.lo[d] = l;
        if (d < NA) {
            const V3 k = ld3(W.jk[d]), o = ld3(W.jo[d]);
            const V3 lv = cross(o - O, k);
            Bd[s][0] = k.x; Bd[s][1] = k.y; Bd[s][2] = k.z; Bd[s][3] = lv.x; Bd[s][4] = lv.y; Bd[s][5] = lv.z;
        } else if (d < NU) {
            const int f = d - NA;
            const V3 k = ld3(W.fk[f]), o = ld3(W.fo[f]);
            if (FH) {
                const V3 lv = cross(o - O, k);
                Bd[s][0] = k.x; Bd[s][1] = k.y; Bd[s][2] = k.z; Bd[s][3] = lv.x; Bd[s][4] = lv.y; Bd[s][5] = lv.z;
            } else {
                Bd[s][3] = k.x; Bd[s][4] = k.y; Bd[s][5] = k.z;
            }
        } else {
            const int kb = (d - NU) / 6, j = (d - NU) % 6;
            const float* o = In.obj[kb];
            float Mc[6], asc[6], qdc[6];
            cube_free_acc(M, mk(o[7], o[8], o[9]), mk(o[10], o[11], o[12]), Mc, asc, qdc);
            if (j < 3) Bd[s][3 + j] = 1.f;
            else {
                const M3 R = qmat(mkq(o[3], o[4], o[5], o[6]));
                const V3 ax = j == 3 ? R.c0 : (j == 4 ? R.c1 : R.c2);
                const V3 lv = cross(mk(o[0], o[1], o[2]) - O, ax);
                Bd[s][0] = ax.x; Bd[s][1] = ax.y; Bd[s][2] = ax.z; Bd[s][3] = lv.x; Bd[s][4] = lv.y; Bd[s][5] = lv.z;
            }
#pragma unroll
            for (int k = 0; k < 6; k++) if (k == j) { as_d[s] = asc[k]; qd_d[s] = qdc[k]; Mc_d[s] = Mc[k]; }
            warm_d[s] = In.warm_obj[kb][j];
        }
        if (d < NU) {                                   // only reached with the robot active, in slot 0
            as_d[s] = W.as[d]; qd_d[s] = W.qd[d]; warm_d[s] = W.warm[d];
            lsg_d = W.lsg[d]; laref_d = W.laref[d]; lD_d = W.lD[d];
#pragma unroll
            for (int j = 0; j < NU; j++) Mrow[j] = W.Mt[d >= j ? GG_TRI(d, j) : GG_TRI(j, d)];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) B.TB[d][k] = Bd[s][k];
    }
    // compact list of the body pairs in contact
    int npair = 0;
    wc.sync();
    for (int p0 = 0; p0 < BufT::NPAIR; p0 += 32) {
        const int pid = p0 + lane;
        const int flag = (pid < BufT::NPAIR && B.pairAB[pid] >= 0) ? 1 : 0;
        const int incl = wc.scan_incl(flag);
        if (flag) B.plist[npair + incl - 1] = pid;
        npair += wc.shfli(incl, 31);
    }
    wc.sync();

    // (Mt + Mc) x for a vector held one component per owned row
    auto mass_mul = [&](const float* x, float* r) {
        if (cpl) {
            wc.sync();
            if (lane < NU) B.vx[lane] = x[0];              // robot rows are local rows 0..NU-1 of slot 0
            wc.sync();
        }
#pragma unroll
        for (int s = 0; s < ND; s++) {
            float v = Mc_d[s] * x[s];
            if (s == 0 && cpl && lane < NU) {
#pragma unroll
                for (int j = 0; j < NU; j++) v += Mrow[j] * B.vx[j];
            }
            r[s] = val[s] ? v : 0.f;
        }
    };
    // body twists of the dof vector x into T[slot] (inactive bodies: zero twist, they carry no contact)
    auto body_twists = [&](const float* x, int slot) {
        wc.sync();
#pragma unroll
        for (int s = 0; s < ND; s++) {
            if (s > 0 && !wide) continue;
            if (val[s]) {
#pragma unroll
                for (int k = 0; k < 6; k++) B.Y[dd[s]][k] = Bd[s][k] * x[s];
            }
        }
        wc.sync();
        for (int t = lane; t < NB * 6; t += 32) {
            const int b = t / 6, k = t % 6;
            float v = 0.f;
            if (b >= 1 && b < 4) {
                if (cpl) {
                    for (int d = 0; d < NA; d++) v += B.Y[d][k];
                    if (b >= 2 && b - 2 < NF) v += B.Y[NA + b - 2][k];
                }
            } else if (b >= 4 && ((amask >> (1 + b - 4)) & 1u)) {
                for (int j = 0; j < 6; j++) v += B.Y[NU + 6 * (b - 4) + j][k];
            }
            B.T[slot][b][k] = v;
        }
        wc.sync();
    };
    // relative velocity of a contact's pair along (n, t1, t2)
    auto rel3 = [&](int c, int slot, float* out3) {
        const float* f = B.F[c];
        const int code = B.code[c];
        const int A = (code >> 8) & 0xff, Bb = code >> 16;
        float r[6];
#pragma unroll
        for (int k = 0; k < 6; k++) r[k] = B.T[slot][Bb][k] - B.T[slot][A][k];
        out3[0] = sk_dot6(f, r); out3[1] = sk_dot6(f + 6, r); out3[2] = sk_dot6(f + 12, r);
    };

    body_twists(qd_d, 0);
    body_twists(warm_d, 1);
    body_twists(as_d, 2);

    // ---- contact rows: reference accelerations, regularisation, residuals for both start candidates
    float Dk[SK_CPL], x[SK_CPL][4], jv[SK_CPL][4];
    float costw = 0.f, costs = 0.f;
#pragma unroll
    for (int s = 0; s < SK_CPL; s++) {
        Dk[s] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) { x[s][k] = 0.f; jv[s][k] = 0.f; }
        const int c = lane + 32 * s;
        if (c < ncon) {
            float vel[3], uw[3], us[3];
            rel3(c, 0, vel); rel3(c, 1, uw); rel3(c, 2, us);
            const int code = B.code[c];
            const int A = (code >> 8) & 0xff, Bb = code >> 16;
            const float wsum = (A == 0 ? 0.f : (A < 4 ? M.body_invweight[A - 1] : M.cube_invweight)) +
                               (Bb < 4 ? M.body_invweight[Bb - 1] : M.cube_invweight);
            float imp = 1.f, ar[4];
            imp_aref4(M, -B.depth[c], vel, mu, imp, ar);
            Dk[s] = gg_rcp(fmaxf(M.minval, M.reg_scale * (1.f - imp) / imp * wsum));
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float jw = rowval(uw, mu, k) - ar[k], js = rowval(us, mu, k) - ar[k];
                if (jw < 0.f) costw += 0.5f * Dk[s] * jw * jw;
                if (js < 0.f) costs += 0.5f * Dk[s] * js * js;
                x[s][k] = jw; jv[s][k] = js;
            }
        }
    }
    {
        float e[ND], Me[ND];
#pragma unroll
        for (int s = 0; s < ND; s++) e[s] = warm_d[s] - as_d[s];
        mass_mul(e, Me);
#pragma unroll
        for (int s = 0; s < ND; s++) costw += 0.5f * e[s] * Me[s];
        const float jw = lsg_d * warm_d[0] - laref_d, js = lsg_d * as_d[0] - laref_d;
        if (lsg_d != 0.f && jw < 0.f) costw += 0.5f * lD_d * jw * jw;
        if (lsg_d != 0.f && js < 0.f) costs += 0.5f * lD_d * js * js;
    }
    costw = wc.sum(costw); costs = wc.sum(costs);
    const bool warm = costw < costs;
    float a_d[ND];
#pragma unroll
    for (int s = 0; s < ND; s++) a_d[s] = warm ? warm_d[s] : as_d[s];
    if (!warm) {
#pragma unroll
        for (int s = 0; s < SK_CPL; s++)
#pragma unroll
            for (int k = 0; k < 4; k++) x[s][k] = jv[s][k];
    }

    // ---- contacts grouped by body pair: perm lists the contact indices pair after pair (contact order within a pair)
    {
        const unsigned lt = (1u << lane) - 1u;
        int run = 0;
        for (int u = 0; u < npair; u++) {
            const int pid = B.plist[u];
            int base = run;
            for (int s = 0; s < SK_CPL; s++) {
                if (32 * s >= ncon) break;                       // uniform
                const int c = lane + 32 * s;
                const bool mine = c < ncon && (B.code[c] & 0xff) == pid;
                const unsigned m = wc.ballot(mine);
                if (mine) B.perm[base + popc32(m & lt)] = c;
                base += popc32(m);
            }
            if (lane == 0) { B.pstart[pid] = run; B.pcnt[pid] = base - run; }
            run = base;
        }
        wc.sync();
    }
    // local (row, column) of the packed entries this lane owns in the projection of a pair onto its <= NA + 1 + 6 dofs
    static_assert((NA + 1 + 6) * (NA + 1 + 6 + 1) / 2 <= 128, "four packed entries per lane cover a body pair");
    int eli[4], elj[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int e = lane + 32 * k;
        eli[k] = sk_tri_row(e); elj[k] = e - eli[k] * (eli[k] + 1) / 2;
    }
    // pair-stiffness lanes: row pi of (Phi | gF), contact subset pg (lanes 30, 31 idle); entry (ei, ej) for the final sum
    const int pi = lane % 6, pg = lane / 6;
    int ei = 0, ej = 0;
    if (lane < 21) {
        ei = lane < 1 ? 0 : (lane < 3 ? 1 : (lane < 6 ? 2 : (lane < 10 ? 3 : (lane < 15 ? 4 : 5))));
        ej = lane - ei * (ei + 1) / 2;
    } else if (lane < 27) { ei = lane - 21; ej = 6; }

    int it = 0;
    for (; it < iterations; it++) {
        // ---- gradient of the smooth part and of the limit rows
        float g_d[ND], hlim = 0.f;
        {
            float e[ND];
#pragma unroll
            for (int s = 0; s < ND; s++) e[s] = a_d[s] - as_d[s];
            mass_mul(e, g_d);
            const float jar = lsg_d * a_d[0] - laref_d;
            if (lsg_d != 0.f && jar < 0.f) { hlim = lD_d; g_d[0] += lD_d * jar * lsg_d; }
        }
        // ---- per-contact weights, clear the accumulators of the pairs in contact
        for (int i = lane; i < n * (n + 1) / 2; i += 32) B.H[i] = 0.f;
#pragma unroll
        for (int s = 0; s < SK_CPL; s++) {
            const int c = lane + 32 * s;
            if (c < ncon) {
                float Wnn = 0.f, Wn1 = 0.f, W11 = 0.f, Wn2 = 0.f, W22 = 0.f, gn = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float jar = x[s][k];
                    if (jar < 0.f) {
                        const float sm = (k & 1) ? mu : -mu;
                        Wnn += Dk[s]; gn += Dk[s] * jar;
                        if (k < 2) { Wn1 += Dk[s] * sm; W11 += Dk[s] * mu * mu; g1 += Dk[s] * jar * sm; }
                        else { Wn2 += Dk[s] * sm; W22 += Dk[s] * mu * mu; g2 += Dk[s] * jar * sm; }
                    }
                }
                B.Wc[0][c] = Wnn; B.Wc[1][c] = Wn1; B.Wc[2][c] = W11; B.Wc[3][c] = Wn2; B.Wc[4][c] = W22;
                B.Wc[5][c] = gn; B.Wc[6][c] = g1; B.Wc[7][c] = g2;
            }
        }
        wc.sync();
        // ---- pair stiffness in spatial-force space, Phi = sum_c F_c^T W_c F_c, gF = sum_c F_c^T g_c per body pair: lane (pi, pg) sums
        //      row pi over the contacts pg, pg + 5, ... of the pair; the five partials are added in a fixed order (bit-reproducible)
        for (int u = 0; u < npair; u++) {
            const int pid = B.plist[u];
            const int s0 = B.pstart[pid], n_c = B.pcnt[pid];
            if (pg < 5) {
                float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accg = 0.f;
                for (int k = pg; k < n_c; k += 5) {
                    const int c = B.perm[s0 + k];
                    const float w0 = B.Wc[0][c], w1 = B.Wc[1][c], w2 = B.Wc[2][c], w3 = B.Wc[3][c], w4 = B.Wc[4][c];
                    const float* f = B.F[c];
                    const float fni = f[pi], f1i = f[6 + pi], f2i = f[12 + pi];
                    const float a = w0 * fni + w1 * f1i + w3 * f2i, b = w1 * fni + w2 * f1i, cc = w3 * fni + w4 * f2i;
#pragma unroll
                    for (int j = 0; j < 6; j++) acc[j] += a * f[j] + b * f[6 + j] + cc * f[12 + j];
                    accg += fni * B.Wc[5][c] + f1i * B.Wc[6][c] + f2i * B.Wc[7][c];
                }
#pragma unroll
                for (int j = 0; j < 6; j++) B.part[pg * 42 + pi * 7 + j] = acc[j];
                B.part[pg * 42 + pi * 7 + 6] = accg;
            }
            wc.sync();
            if (lane < 27) {
                const int idx = ei * 7 + ej;
                const float v = (((B.part[idx] + B.part[42 + idx]) + B.part[84 + idx]) + B.part[126 + idx]) + B.part[168 + idx];
                if (lane < 21) B.Phi[pid][lane] = v; else B.gF[pid][lane - 21] = v;
            }
            wc.sync();
        }
        // ---- H = Mt (robot block, cube diagonal) ...
        if (cpl) for (int e = lane; e < NU * (NU + 1) / 2; e += 32) B.H[e] = W.Mt[e];
#pragma unroll
        for (int s = 0; s < ND; s++) {
            const int l = lane + 32 * s;
            if (val[s] && dd[s] >= NU) B.H[GG_TRI(l, l)] = Mc_d[s];
        }
        wc.sync();
        // ---- ... + sum over the body pairs in contact of P^T Phi P;  g += P^T gF
        for (int u = 0; u < npair; u++) {
            const int pid = B.plist[u];
            const int ab = B.pairAB[pid];
            const int A = ab & 0xff, Bb = ab >> 8;
#pragma unroll
            for (int s = 0; s < ND; s++) {
                if (s > 0 && !wide) continue;
                if (val[s]) {
                    const int d = dd[s];
                    const float sj = sk_sign<NA, NF>(A, Bb, d);
                    if (sj != 0.f) {
                        const float* P = B.Phi[pid];
#pragma unroll
                        for (int r = 0; r < 6; r++) {
                            float v = 0.f;
#pragma unroll
                            for (int c = 0; c < 6; c++) v += P[r >= c ? GG_TRI(r, c) : GG_TRI(c, r)] * Bd[s][c];
                            B.Y[d][r] = v * sj;
                        }
                        g_d[s] += sj * sk_dot6(Bd[s], B.gF[pid]);
                    }
                }
            }
            wc.sync();
            const int nA = sk_ndof<NA, NF>(A), nl = nA + sk_ndof<NA, NF>(Bb);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int e = lane + 32 * k;
                if (e >= nl * (nl + 1) / 2) continue;
                const int li = eli[k], lj = elj[k];
                const int i = li < nA ? sk_dof<NA, NF>(A, li) : sk_dof<NA, NF>(Bb, li - nA);
                const int j = lj < nA ? sk_dof<NA, NF>(A, lj) : sk_dof<NA, NF>(Bb, lj - nA);
                const float si = sk_sign<NA, NF>(A, Bb, i);
                B.H[GG_TRI(B.lo[i], B.lo[j])] += si * sk_dot6(B.TB[i], B.Y[j]);
            }
            wc.sync();
        }
        if (cpl && lane < NU && hlim != 0.f) B.H[GG_TRI(lane, lane)] += hlim;
        wc.sync();
        // ---- stop criterion: max_i |g_i| / H_ii <= tol (1 + |a|_inf)
        {
            float r = 0.f, am = 0.f;
#pragma unroll
            for (int s = 0; s < ND; s++) {
                const int l = lane + 32 * s;
                if (val[s]) { r = fmaxf(r, fabsf(g_d[s]) * gg_rcp(B.H[GG_TRI(l, l)])); am = fmaxf(am, fabsf(a_d[s])); }
            }
            r = wc.max(r); am = wc.max(am);
            if (gg_div(r, 1.f + am) < tol) break;
        }
        // ---- Cholesky in place (left-looking, lane = row): L(i, j) = (H(i, j) - sum_{k<j} L(i, k) L(j, k)) / L(j, j)
        bool ok = true;
        for (int j = 0; j < n; j++) {
            float t[ND];
#pragma unroll
            for (int s = 0; s < ND; s++) {
                const int i = j + lane + 32 * s;
                t[s] = 0.f;
                if (s > 0 && n - j <= 32) continue;
                if (i < n) {
                    const float* ri = &B.H[GG_TRI(i, 0)];
                    const float* rj = &B.H[GG_TRI(j, 0)];
                    float v = ri[j];
#pragma unroll 4
                    for (int k = 0; k < j; k++) v -= ri[k] * rj[k];
                    t[s] = v;
                }
            }
            const float piv = wc.shfl(t[0], 0);
            ok = ok && (piv > 0.f);
            const float inv = gg_rsqrt(piv);
#pragma unroll
            for (int s = 0; s < ND; s++) {
                const int i = j + lane + 32 * s;
                if (i < n) B.H[GG_TRI(i, j)] = i == j ? piv * inv : t[s] * inv;
            }
            if (lane == 0) B.dinv[j] = inv;
            wc.sync();
        }
        if (!ok) break;
        // ---- solve L L^T s = -g: the right-hand side stays in registers (lane owns its rows), one shuffle per column
        float s_d[ND];
#pragma unroll
        for (int s = 0; s < ND; s++) s_d[s] = val[s] ? -g_d[s] : 0.f;
        for (int k = 0; k < n; k++) {
            float cand = s_d[0];
#pragma unroll
            for (int s = 1; s < ND; s++) if ((k >> 5) == s) cand = s_d[s];
            const float yk = wc.shfl(cand, k & 31) * B.dinv[k];
#pragma unroll
            for (int s = 0; s < ND; s++) {
                const int l = lane + 32 * s;
                if (l == k) s_d[s] = yk;
                else if (l > k && l < n) s_d[s] -= B.H[GG_TRI(l, k)] * yk;
            }
        }
        for (int k = n - 1; k >= 0; k--) {
            float cand = s_d[0];
#pragma unroll
            for (int s = 1; s < ND; s++) if ((k >> 5) == s) cand = s_d[s];
            const float xk = wc.shfl(cand, k & 31) * B.dinv[k];
#pragma unroll
            for (int s = 0; s < ND; s++) {
                const int l = lane + 32 * s;
                if (l == k) s_d[s] = xk;
                else if (l < k) s_d[s] -= B.H[GG_TRI(k, l)] * xk;
            }
        }
        // ---- exact line search along s
        float q1 = 0.f, q2 = 0.f;
        {
            float Ms[ND];
            mass_mul(s_d, Ms);
#pragma unroll
            for (int s = 0; s < ND; s++) { q1 += Ms[s] * (a_d[s] - as_d[s]); q2 += Ms[s] * s_d[s]; }
            q1 = wc.sum(q1); q2 = wc.sum(q2);
        }
        body_twists(s_d, 2);
#pragma unroll
        for (int s = 0; s < SK_CPL; s++) {
            const int c = lane + 32 * s;
            if (c < ncon) {
                float w3[3];
                rel3(c, 2, w3);
#pragma unroll
                for (int k = 0; k < 4; k++) jv[s][k] = rowval(w3, mu, k);
            }
        }
        LineSearch Ls; Ls.begin();
        const float ljv = lsg_d * s_d[0], lx0 = lsg_d * a_d[0] - laref_d;
        for (;;) {
            float d1 = 0.f, d2 = 0.f;
            {
                const float xx = lx0 + Ls.al * ljv;
                if (lsg_d != 0.f && xx < 0.f) { d1 += lD_d * xx * ljv; d2 += lD_d * ljv * ljv; }
            }
#pragma unroll
            for (int s = 0; s < SK_CPL; s++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float xx = x[s][k] + Ls.al * jv[s][k];
                    if (xx < 0.f) { d1 += Dk[s] * xx * jv[s][k]; d2 += Dk[s] * jv[s][k] * jv[s][k]; }
                }
            d1 = wc.sum(d1) + q1 + Ls.al * q2; d2 = wc.sum(d2) + q2;
            if (!Ls.next(d1, d2)) break;
        }
        const float al = Ls.al;
        bool changed = (lsg_d != 0.f) && ((lx0 < 0.f) != (lx0 + al * ljv < 0.f));
#pragma unroll
        for (int s = 0; s < ND; s++) a_d[s] += al * s_d[s];
#pragma unroll
        for (int s = 0; s < SK_CPL; s++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float x1 = x[s][k] + al * jv[s][k];
                changed = changed || ((Dk[s] != 0.f) && ((x[s][k] < 0.f) != (x1 < 0.f)));
                x[s][k] = x1;
            }
        if (!wc.any(changed) && fabsf(al - 1.f) < 1e-3f) { it++; break; }
    }
    iters_out = it;
#pragma unroll
    for (int s = 0; s < ND; s++)
        if (val[s]) a_out[(size_t)dd[s] * a_stride] = a_d[s];
    wc.sync();
}

// ---------------------------------------------------------------------------------------------------------------------
// stage 3: thread per env
// ---------------------------------------------------------------------------------------------------------------------
// Observation / reward / success of the stacking task (franka/cube_stack.py:188-216): agent 7 + NF floats (or the n_dof joint
// positions, T.agent_qpos), envs 14 floats.
// The reward compares float32 device tensors: sqrt(dx^2 + dy^2) < 0.05f and z1 - z2 > 0.03f; the 2-norm is evaluated as
// sqrtf(fmaf(dx, dx, dy * dy)) here and in the oracle so that the flags are bit-exact given matched states.
template <int NA, int NF, int K>
GG_HD void stack_observe(const GgModel& M, const GgStackTask& T, const float* q, const ObjState<K>& O, float* agent, float* envs,
                         float& reward, bool& success) {
    V3 ep; Q4 eq;
    robot_eef<NA, NF>(M, q, ep, eq);
    if (T.agent_qpos) {
        // so101/cube_stack_batch.py:176: agent_pos = so_101.get_qpos()  (n_dof floats; the buffer holds 7 + NF >= n_dof)
#pragma unroll
        for (int i = 0; i < NA + NF; i++) agent[i] = q[i];
    } else {
        agent[0] = ep.x; agent[1] = ep.y; agent[2] = ep.z;
        agent[3] = eq.w; agent[4] = eq.x; agent[5] = eq.y; agent[6] = eq.z;
#pragma unroll
        for (int f = 0; f < NF; f++) agent[7 + f] = q[NA + f];
    }
    const V3 c1 = O.p[0], c2 = O.p[K > 1 ? 1 : 0];
    envs[0] = c1.x; envs[1] = c1.y; envs[2] = c1.z;
    envs[3] = O.q[0].w; envs[4] = O.q[0].x; envs[5] = O.q[0].y; envs[6] = O.q[0].z;
    const V3 d = ep - c1;
    envs[7] = d.x; envs[8] = d.y; envs[9] = d.z;
    envs[10] = sqrtf(dot(d, d));
    envs[11] = c2.x; envs[12] = c2.y; envs[13] = c2.z;
    const float dx = c1.x - c2.x, dy = c1.y - c2.y;
#if defined(__CUDA_ARCH__)
    const float xy = sqrtf(fmaf(dx, dx, __fmul_rn(dy, dy)));
#else
    const float xy = sqrtf(fmaf(dx, dx, dy * dy));
#endif
    const float dz = c1.z - c2.z;
    success = xy < T.stack_xy && dz > T.stack_z;
    reward = success ? 1.f : 0.f;
}

// device-side spawn draw of object k: Philox counter (global env id, episode, k, attempt), key = seed
GG_HD void stack_spawn_draw(const GgStackTask& T, uint64_t seed, uint32_t gid, uint32_t episode, int k, uint32_t attempt, float& x, float& y) {
    uint32_t o[4];
    philox4x32_10(gid, episode, (uint32_t)k, attempt, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    x = fmaf(u01(o[0]), T.spawn_hi[k][0] - T.spawn_lo[k][0], T.spawn_lo[k][0]);
    y = fmaf(u01(o[1]), T.spawn_hi[k][1] - T.spawn_lo[k][1], T.spawn_lo[k][1]);
}
// all objects of one env.  With T.min_dist > 0 objects 0 and 1 are re-drawn TOGETHER (attempt 0, 1, ... <= 15) until
// fmaf(dx, dx, dy * dy) >= min_dist^2 -- the rejection loop of so101/cube_stack_batch.py:70-86 on the device stream.
GG_HD void stack_spawn_all(const GgStackTask& T, uint64_t seed, uint32_t gid, uint32_t episode, int K, float (*xy)[2]) {
    for (int k = 0; k < K; k++) stack_spawn_draw(T, seed, gid, episode, k, 0u, xy[k][0], xy[k][1]);
    if (T.min_dist > 0.f && K > 1) {
        for (uint32_t a = 1; a < 16; a++) {
            const float dx = xy[1][0] - xy[0][0], dy = xy[1][1] - xy[0][1];
#if defined(__CUDA_ARCH__)
            const float d2 = fmaf(dx, dx, __fmul_rn(dy, dy));
#else
            const float d2 = fmaf(dx, dx, dy * dy);
#endif
            if (d2 >= T.min_dist * T.min_dist) break;
            stack_spawn_draw(T, seed, gid, episode, 0, a, xy[0][0], xy[0][1]);
            stack_spawn_draw(T, seed, gid, episode, 1, a, xy[1][0], xy[1][1]);
        }
    }
}

}  // namespace gg
