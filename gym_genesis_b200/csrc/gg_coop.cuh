// Warp-cooperative coupled solve (device only): ONE WARP works on ONE env whose gripper pads may touch the cube or the
// plane.  Same algorithm as solve_coupled in gg_env.cuh (narrow phase in the same order, same soft-constraint cost,
// Newton with exact line search, same stop criteria) mapped onto 32 lanes:
//   * narrow phase          lane = pad box (plane-box and box-box run concurrently for all pads), ordered compaction
//   * per-contact rows      lane = contact (2 per lane at the 64-contact cap): residuals x = J a - aref live in registers
//   * pair stiffness        lane = (row i of (Phi, gF) of a body pair, one of five contact subsets); fixed-order sum of the five partials
//   * projection            lane = dof for Y = Phi B_j, then lane = Hessian entry (4 per lane) for H += B_i . Y_j
//   * Cholesky / solves     lane = row of the 15 x 15 factor kept in registers, columns broadcast with shuffles
//   * line search           every lane evaluates its own limit row / contacts, butterfly reduction of (phi', phi'')
// Working set lives in shared memory (TeamBuf, ~10 KB per warp) and registers: no local-memory arrays, no divergence
// between envs of a warp (the thread-per-env solver spilled 8.9 KB / thread to DRAM and ran every env of a warp for the
// slowest env's iteration count: profiles/r01_notes.md, v4 grasp capture).
#pragma once
#include "gg_env.cuh"

namespace gg {

constexpr unsigned kFull = 0xffffffffu;

template <int NA, int NF> struct TeamWork {          // published by the MAIN lane of the env (shared memory)
    static constexpr int NU = NA + NF;
    float Mt[NU * (NU + 1) / 2];
    float as[NU], qd[NU], warm[NU], lsg[NU], laref[NU], lD[NU];
    float jk[NA][3], jo[NA][3];
    float fk[NF > 0 ? NF : 1][3], fo[NF > 0 ? NF : 1][3];
    float lR[9], lp[3];
    float fR[NF > 0 ? NF : 1][9], fp[NF > 0 ? NF : 1][3];
    float cube[19];                                   // cp3 cq4 cv3 cw3 qacc6
    float q[NU];                                      // joint positions (the drain warp of step_rs_kernel integrates the env itself)
};

#ifndef GG_TEAM_CAP
#define GG_TEAM_CAP GG_MAX_CONTACTS
#endif
template <int NV> struct TeamBuf {
    static constexpr int NH = NV * (NV + 1) / 2;
    float F[GG_TEAM_CAP][18];     // spatial forces (moment about O, direction) of n, t1, t2; first p, n while colliding
    float Wc[8][GG_TEAM_CAP];     // per-iteration weights Wnn Wn1 W11 Wn2 W22 gn g1 g2, component-major: the writer is lane = contact
                                  // (a [contact][8] layout is an 8-way bank conflict), the readers broadcast one contact
    float depth[GG_MAX_CONTACTS];
    int code[GG_MAX_CONTACTS];
    int perm[GG_MAX_CONTACTS];        // contact indices grouped by body pair (stable), see team_solve
    float part[5 * 42];               // partial pair-stiffness sums of the five contact subsets
    float TB[NV][6];                  // twist of every dof about O (angular, linear)
    float T[3][7][6];                 // relative pair twists for qd / warm start / a_smooth (slot 2 reused for the search direction)
    float Phi[7][21], gF[7][6];
    float Y[NV][6];
    float H[NH];
    float L[NH];
    float vec[32];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// sign with which dof d enters the relative twist of pair pc (0: not involved).
// pairs: 0 plane-cube, 1 + r plane-robot body r, 4 + r robot body r-cube (r: 0 last arm link, 1 + f finger f)
template <int NA, int NF> __device__ __forceinline__ float pair_sign(int pc, int d) {
    constexpr int NU = NA + NF;
    if (pc == 0) return d >= NU ? 1.f : 0.f;
    const int rb = (pc - 1) % 3;
    const float sr = pc >= 4 ? -1.f : 1.f;
    if (d < NA) return sr;
    if (d < NU) return (rb >= 1 && d - NA == rb - 1) ? sr : 0.f;
    return pc >= 4 ? 1.f : 0.f;
}

__device__ __forceinline__ float dot6(const float* a, const float* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// relative pair twists T[pair][6] of the dof vector whose component d sits in lane d (x), via the scratch Y
template <int NA, int NF, int NV>
__device__ __forceinline__ void team_pair_twists(TeamBuf<NV>& B, const float* Bd, float x, int lane, float (*T)[6]) {
    constexpr int NU = NA + NF;
    if (lane < NV) {
#pragma unroll
        for (int k = 0; k < 6; k++) B.Y[lane][k] = Bd[k] * x;
    }
    __syncwarp();
    if (lane < 6) {
        float arm = 0.f, cube = 0.f;
#pragma unroll
        for (int d = 0; d < NA; d++) arm += B.Y[d][lane];
#pragma unroll
        for (int d = NU; d < NV; d++) cube += B.Y[d][lane];
        const float f0 = NF > 0 ? B.Y[NA][lane] : 0.f;
        const float f1 = NF > 1 ? B.Y[NA + (NF > 1 ? 1 : 0)][lane] : 0.f;
        T[0][lane] = cube; T[1][lane] = arm; T[2][lane] = arm + f0; T[3][lane] = arm + f1;
        T[4][lane] = cube - arm; T[5][lane] = cube - (arm + f0); T[6][lane] = cube - (arm + f1);
    }
    __syncwarp();
}

// ---- narrow phase, one pad per lane; contacts come out in the serial order (cube-plane, pads-plane, pads-cube) ------
template <int NA, int NF>
__device__ __forceinline__ int team_collide(const GgModel& M, const TeamWork<NA, NF>& W, V3 cp, const M3& cR, TeamBuf<NA + NF + 6>& B, int lane) {
    const int maxc = M.max_contacts < GG_MAX_CONTACTS ? M.max_contacts : GG_MAX_CONTACTS;
    const V3 ch = ld3(M.cube_half);
    int total = 0;
    // pass -1: cube-plane (lane 0); pass 0: pads-plane; pass 1: pads-cube  (lane g = pad g)
#pragma unroll 1
    for (int pass = -1; pass < 2; pass++) {
        Contacts8 C;                  // per-lane mini list
        C.count = 0;
        if (pass < 0) {
            if (lane == 0) plane_box(M.plane_z, cp, cR, ch, 0, 8, C);
        } else if (lane < M.n_pads) {
            const int g = lane;
            const int b = M.pad_body[g];
            const int fi = b - 1 < 0 ? 0 : (b - 1 < NF ? b - 1 : 0);
            M3 Rb; V3 pb;
            if (b == 0) {
                Rb.c0 = mk(W.lR[0], W.lR[1], W.lR[2]); Rb.c1 = mk(W.lR[3], W.lR[4], W.lR[5]); Rb.c2 = mk(W.lR[6], W.lR[7], W.lR[8]);
                pb = mk(W.lp[0], W.lp[1], W.lp[2]);
            } else {
                const float* r = W.fR[fi];
                Rb.c0 = mk(r[0], r[1], r[2]); Rb.c1 = mk(r[3], r[4], r[5]); Rb.c2 = mk(r[6], r[7], r[8]);
                pb = mk(W.fp[fi][0], W.fp[fi][1], W.fp[fi][2]);
            }
            const V3 c = pb + mul(Rb, ld3(M.pad_pos[g]));
            const M3 R = mmul(Rb, qmat(ldq(M.pad_quat[g])));
            const V3 h = ld3(M.pad_half[g]);
            if (pass == 0) plane_box(M.plane_z, c, R, h, 1 + b, 8, C);
            else box_box(c, R, h, cp, cR, ch, 4 + b, 8, C);
        }
        // ordered compaction: exclusive prefix of the per-lane counts
        int incl = C.count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const int base = total + incl - C.count;
        for (int k = 0; k < C.count; k++) {
            const int c = base + k;
            if (c < maxc) {
                B.F[c][0] = C.p[k].x; B.F[c][1] = C.p[k].y; B.F[c][2] = C.p[k].z;
                B.F[c][3] = C.n[k].x; B.F[c][4] = C.n[k].y; B.F[c][5] = C.n[k].z;
                B.depth[c] = C.depth[k]; B.code[c] = C.code[k];
            }
        }
        total += __shfl_sync(kFull, incl, 31);
        if (total > maxc) total = maxc;
    }
    __syncwarp();
    return total;
}

// ---- the solve.  On return lane d < NV holds a[d] (also written to a_out[d * a_stride]). -----------------------------
template <int NA, int NF, bool FH>
__device__ __noinline__ void team_solve(const GgModel& M, const TeamWork<NA, NF>& W, TeamBuf<NA + NF + 6>& B, int iterations, float tol,
                                           int lane, float* a_out, int a_stride, int& iters_out, int& ncon_out) {
    constexpr int NU = NA + NF, NV = NU + 6, NH = NV * (NV + 1) / 2;
    const float mu = M.friction;
    const V3 cp = mk(W.cube[0], W.cube[1], W.cube[2]);
    const Q4 cq = mkq(W.cube[3], W.cube[4], W.cube[5], W.cube[6]);
    const V3 cv = mk(W.cube[7], W.cube[8], W.cube[9]), cw = mk(W.cube[10], W.cube[11], W.cube[12]);
    const M3 cR = qmat(cq);
    const V3 O = cp;

    // ---- contacts
    const int ncon = team_collide<NA, NF>(M, W, cp, cR, B, lane);
    ncon_out = ncon;
    unsigned used = 0u;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int c = lane + 32 * s;
        if (c < ncon) {
            const V3 p = mk(B.F[c][0], B.F[c][1], B.F[c][2]) - O, n = mk(B.F[c][3], B.F[c][4], B.F[c][5]);
            V3 t1, t2;
            tangent_basis(n, t1, t2);
            const V3 pn = cross(p, n), p1 = cross(p, t1), p2 = cross(p, t2);
            float* f = B.F[c];
            f[0] = pn.x; f[1] = pn.y; f[2] = pn.z; f[3] = n.x; f[4] = n.y; f[5] = n.z;
            f[6] = p1.x; f[7] = p1.y; f[8] = p1.z; f[9] = t1.x; f[10] = t1.y; f[11] = t1.z;
            f[12] = p2.x; f[13] = p2.y; f[14] = p2.z; f[15] = t2.x; f[16] = t2.y; f[17] = t2.z;
            used |= 1u << B.code[c];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) used |= __shfl_xor_sync(kFull, used, o);

    // ---- per-dof data: lane d < NV owns dof d
    const int d = lane;
    float Bd[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float as_d = 0.f, qd_d = 0.f, warm_d = 0.f, lsg_d = 0.f, laref_d = 0.f, lD_d = 0.f, Mc_d = 0.f;
    float Mrow[NU];
#pragma unroll
    for (int j = 0; j < NU; j++) Mrow[j] = 0.f;
    {
        float Mc[6], asc[6], qdc[6];
        cube_free_acc(M, cv, cw, Mc, asc, qdc);
        if (d < NA) {
            const V3 k = mk(W.jk[d][0], W.jk[d][1], W.jk[d][2]), o = mk(W.jo[d][0], W.jo[d][1], W.jo[d][2]);
            const V3 l = cross(o - O, k);
            Bd[0] = k.x; Bd[1] = k.y; Bd[2] = k.z; Bd[3] = l.x; Bd[4] = l.y; Bd[5] = l.z;
        } else if (d < NU) {
            const int f = d - NA;
            const V3 k = mk(W.fk[f][0], W.fk[f][1], W.fk[f][2]), o = mk(W.fo[f][0], W.fo[f][1], W.fo[f][2]);
            if (FH) {
                const V3 l = cross(o - O, k);
                Bd[0] = k.x; Bd[1] = k.y; Bd[2] = k.z; Bd[3] = l.x; Bd[4] = l.y; Bd[5] = l.z;
            } else {
                Bd[3] = k.x; Bd[4] = k.y; Bd[5] = k.z;
            }
        } else if (d < NV) {
            const int j = d - NU;
            if (j < 3) Bd[3 + j] = 1.f;
            else {
                const V3 c = j == 3 ? cR.c0 : (j == 4 ? cR.c1 : cR.c2);
                Bd[0] = c.x; Bd[1] = c.y; Bd[2] = c.z;
            }
#pragma unroll
            for (int k = 0; k < 6; k++) if (k == j) { as_d = asc[k]; qd_d = qdc[k]; Mc_d = Mc[k]; }
            warm_d = W.cube[13 + j];
        }
        if (d < NU) {
            as_d = W.as[d]; qd_d = W.qd[d]; warm_d = W.warm[d]; lsg_d = W.lsg[d]; laref_d = W.laref[d]; lD_d = W.lD[d];
#pragma unroll
            for (int j = 0; j < NU; j++) Mrow[j] = W.Mt[d >= j ? GG_TRI(d, j) : GG_TRI(j, d)];
        }
        if (d < NV) {
#pragma unroll
            for (int k = 0; k < 6; k++) B.TB[d][k] = Bd[k];
        }
    }
    // (Mt + Mc) x for the vector held one component per lane
    auto mass_mul = [&](float x) -> float {
        float r = Mc_d * x;
#pragma unroll
        for (int j = 0; j < NU; j++) r += Mrow[j] * __shfl_sync(kFull, x, j);
        return r;
    };

    team_pair_twists<NA, NF, NV>(B, Bd, qd_d, lane, B.T[0]);
    team_pair_twists<NA, NF, NV>(B, Bd, warm_d, lane, B.T[1]);
    team_pair_twists<NA, NF, NV>(B, Bd, as_d, lane, B.T[2]);

    // ---- contact rows: reference accelerations, regularisation, residuals for both start candidates
    float Dk[2] = {0.f, 0.f}, x[2][4], jv[2][4];
    float costw = 0.f, costs = 0.f;
#pragma unroll
    for (int s = 0; s < 2; s++) {
#pragma unroll
        for (int k = 0; k < 4; k++) { x[s][k] = 0.f; jv[s][k] = 0.f; }
        const int c = lane + 32 * s;
        if (c < ncon) {
            const float* f = B.F[c];
            const int code = B.code[c];
            const float vel[3] = {dot6(f, B.T[0][code]), dot6(f + 6, B.T[0][code]), dot6(f + 12, B.T[0][code])};
            const float uw[3] = {dot6(f, B.T[1][code]), dot6(f + 6, B.T[1][code]), dot6(f + 12, B.T[1][code])};
            const float us[3] = {dot6(f, B.T[2][code]), dot6(f + 6, B.T[2][code]), dot6(f + 12, B.T[2][code])};
            const int rb = code == 0 ? -1 : (code - 1) % 3;
            const float wsum = (rb >= 0 ? M.body_invweight[rb] : 0.f) + ((code == 0 || code >= 4) ? M.cube_invweight : 0.f);
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
        const float e = warm_d - as_d;
        const float Me = mass_mul(e);
        costw += 0.5f * e * Me;
        const float jw = lsg_d * warm_d - laref_d, js = lsg_d * as_d - laref_d;
        if (lsg_d != 0.f && jw < 0.f) costw += 0.5f * lD_d * jw * jw;
        if (lsg_d != 0.f && js < 0.f) costs += 0.5f * lD_d * js * js;
    }
    costw = warp_sum(costw); costs = warp_sum(costs);
    const bool warm = costw < costs;
    float a_d = warm ? warm_d : as_d;
    if (!warm) {
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int k = 0; k < 4; k++) x[s][k] = jv[s][k];
    }

    // ---- contacts grouped by body pair (stable counting sort; the narrow phase emits them pair after pair already unless the
    //      model lists its pads out of body order): perm = contact indices pair after pair, lane pc keeps the range of pair pc
    int pstart = 0, pcnt = 0;
    {
        const int c0 = lane < ncon ? B.code[lane] : 31, c1 = lane + 32 < ncon ? B.code[lane + 32] : 31;
        const unsigned lt = (1u << lane) - 1u;
        int r0 = 0, r1 = 0, run = 0;
#pragma unroll
        for (int pc = 0; pc < 7; pc++) {
            const unsigned b0 = __ballot_sync(kFull, c0 == pc), b1 = __ballot_sync(kFull, c1 == pc);
            const int n0 = __popc(b0), n1 = __popc(b1);
            if (c0 == pc) r0 = run + __popc(b0 & lt);
            if (c1 == pc) r1 = run + n0 + __popc(b1 & lt);
            if (lane == pc) { pstart = run; pcnt = n0 + n1; }
            run += n0 + n1;
        }
        if (lane < ncon) B.perm[r0] = lane;
        if (lane + 32 < ncon) B.perm[r1] = lane + 32;
    }
    // packed Hessian entries owned by this lane (4 per lane) and their constant part (mass matrix / cube inertia)
    float Hbase[4];
    int hi[4], hj[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int e = lane + 32 * k;
        int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        if ((i + 1) * (i + 2) / 2 <= e) i++;                     // guard the float round-off of the closed form
        if (i * (i + 1) / 2 > e) i--;
        hi[k] = i; hj[k] = e - i * (i + 1) / 2;
        float base = 0.f;
        if (e < NH) {
            if (i < NU) base = W.Mt[e];                          // robot block shares the packed index
            else if (hj[k] == i) base = (i - NU < 3) ? M.cube_mass : M.cube_inertia[i - NU - 3];
        }
        Hbase[k] = base;
    }
    // pair signs of this lane's dof and of the rows of its Hessian entries, two bits per pair (bit 2 pc: involved, bit 2 pc + 1: negative)
    unsigned sgn_d = 0u, sgn_h[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int pc = 0; pc < 7; pc++) {
        const float sd = d < NV ? pair_sign<NA, NF>(pc, d) : 0.f;
        sgn_d |= (sd != 0.f ? 1u : 0u) << (2 * pc) | (sd < 0.f ? 2u : 0u) << (2 * pc);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float sh = lane + 32 * k < NH ? pair_sign<NA, NF>(pc, hi[k]) : 0.f;
            sgn_h[k] |= (sh != 0.f ? 1u : 0u) << (2 * pc) | (sh < 0.f ? 2u : 0u) << (2 * pc);
        }
    }
    auto sgn_of = [](unsigned m, int pc) -> float {
        const unsigned b = m >> (2 * pc);
        return (b & 1u) ? ((b & 2u) ? -1.f : 1.f) : 0.f;
    };
    // pair-stiffness lanes: row pi of (Phi | gF), contact subset pg (lanes 30, 31 idle); entry (ei, ej) for the final sum
    const int pi = lane % 6, pg = lane / 6;
    int ei = 0, ej = 0;
    if (lane < 21) {
        ei = lane < 1 ? 0 : (lane < 3 ? 1 : (lane < 6 ? 2 : (lane < 10 ? 3 : (lane < 15 ? 4 : 5))));
        ej = lane - ei * (ei + 1) / 2;
    } else if (lane < 27) { ei = lane - 21; ej = 6; }
    __syncwarp();

    int it = 0;
#pragma unroll 1
    for (; it < iterations; it++) {
        // ---- gradient of the smooth part and of the limit rows (lane = dof)
        float g_d = mass_mul(a_d - as_d);
        float hlim = 0.f;
        {
            const float jar = lsg_d * a_d - laref_d;
            if (lsg_d != 0.f && jar < 0.f) { hlim = lD_d; g_d += lD_d * jar * lsg_d; }
        }
        // ---- stage the per-contact weights
#pragma unroll
        for (int s = 0; s < 2; s++) {
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
        __syncwarp();
        // ---- pair stiffness in spatial-force space, Phi = sum_c F_c^T W_c F_c, gF = sum_c F_c^T g_c per body pair: lane (pi, pg) sums
        //      row pi over the contacts pg, pg + 5, ... of the pair; the five partials are added in a fixed order (bit-reproducible)
        {
            float* part = B.part;
#pragma unroll 1
            for (int pc = 0; pc < 7; pc++) {
                if (!((used >> pc) & 1u)) continue;
                const int s0 = __shfl_sync(kFull, pstart, pc), n = __shfl_sync(kFull, pcnt, pc);
                if (pg < 5) {
                    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accg = 0.f;
                    for (int k = pg; k < n; k += 5) {
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
                    for (int j = 0; j < 6; j++) part[pg * 42 + pi * 7 + j] = acc[j];
                    part[pg * 42 + pi * 7 + 6] = accg;
                }
                __syncwarp();
                if (lane < 27) {
                    const int idx = ei * 7 + ej;
                    const float v = (((part[idx] + part[42 + idx]) + part[84 + idx]) + part[126 + idx]) + part[168 + idx];
                    if (lane < 21) B.Phi[pc][lane] = v; else B.gF[pc][lane - 21] = v;
                }
                __syncwarp();
            }
        }
        // ---- project every used pair onto its dofs: H += P^T Phi P, g += P^T gF
        float Hacc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) Hacc[k] = Hbase[k];
#pragma unroll 1
        for (int pc = 0; pc < 7; pc++) {
            if (!((used >> pc) & 1u)) continue;
            const float sj = sgn_of(sgn_d, pc);
            if (d < NV) {
                float y[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (sj != 0.f) {
                    const float* P = B.Phi[pc];
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        float v = 0.f;
#pragma unroll
                        for (int c = 0; c < 6; c++) v += P[r >= c ? GG_TRI(r, c) : GG_TRI(c, r)] * Bd[c];
                        y[r] = v * sj;
                    }
                    g_d += sj * dot6(Bd, B.gF[pc]);
                }
#pragma unroll
                for (int r = 0; r < 6; r++) B.Y[d][r] = y[r];
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float si = sgn_of(sgn_h[k], pc);
                if (si != 0.f) Hacc[k] += si * dot6(B.TB[hi[k]], B.Y[hj[k]]);
            }
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 4; k++) if (lane + 32 * k < NH) B.H[lane + 32 * k] = Hacc[k];
        __syncwarp();
        // ---- lane i takes row i of H (+ limit stiffness on the diagonal)
        float r[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) r[j] = (d < NV && j <= d) ? B.H[GG_TRI(d < NV ? d : 0, j <= d ? j : 0)] : 0.f;
        float hdd = 1.f;
#pragma unroll
        for (int j = 0; j < NV; j++) if (j == d) { r[j] += hlim; hdd = r[j]; }
        // ---- stop criterion: max_i |g_i| / H_ii <= tol (1 + |a|_inf)
        {
            const float res = warp_max(d < NV ? fabsf(g_d) * gg_rcp(hdd) : 0.f);
            const float am = warp_max(d < NV ? fabsf(a_d) : 0.f);
            if (gg_div(res, 1.f + am) < tol) break;
        }
        // ---- Cholesky, row i in lane i: L(i, j) = (H(i, j) - sum_{k<j} L(i, k) L(j, k)) / L(j, j)
        bool ok = true;
        float dinv_d = 1.f;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            float t = r[j];
#pragma unroll
            for (int k = 0; k < j; k++) t -= r[k] * __shfl_sync(kFull, r[k], j);
            const float piv = __shfl_sync(kFull, t, j);          // lane j's t is the pivot H'(j, j)
            ok = ok && (piv > 0.f);
            const float inv = gg_rsqrt(piv);
            r[j] = d == j ? piv * inv : t * inv;
            if (d == j) dinv_d = inv;
        }
        if (!ok) break;
        if (d < NV) {
#pragma unroll
            for (int j = 0; j < NV; j++) if (j <= d) B.L[GG_TRI(d, j)] = r[j];
        }
        __syncwarp();
        // forward: y_i = (b_i - sum_{k<i} L(i, k) y_k) / L(i, i),  b = -g
        float y_d = -g_d;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const float yk = __shfl_sync(kFull, y_d * dinv_d, k);   // final for lane k at step k
            if (d == k) y_d = yk;
            else if (d > k) y_d -= r[k] * yk;
        }
        // backward: x_i = (y_i - sum_{k>i} L(k, i) x_k) / L(i, i)
        float s_d = y_d;
#pragma unroll
        for (int k = NV - 1; k >= 0; k--) {
            const float xk = __shfl_sync(kFull, s_d * dinv_d, k);
            if (d == k) s_d = xk;
            else if (d < k) s_d -= B.L[GG_TRI(k, d < NV ? d : 0)] * xk;
        }
        if (d >= NV) s_d = 0.f;
        // ---- exact line search along s
        const float Ms = mass_mul(s_d);
        const float q1 = warp_sum(Ms * (a_d - as_d)), q2 = warp_sum(Ms * s_d);
        team_pair_twists<NA, NF, NV>(B, Bd, s_d, lane, B.T[2]);
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int c = lane + 32 * s;
            if (c < ncon) {
                const float* f = B.F[c];
                const int code = B.code[c];
                const float w3[3] = {dot6(f, B.T[2][code]), dot6(f + 6, B.T[2][code]), dot6(f + 12, B.T[2][code])};
#pragma unroll
                for (int k = 0; k < 4; k++) jv[s][k] = rowval(w3, mu, k);
            }
        }
        LineSearch Ls; Ls.begin();
        const float ljv = lsg_d * s_d, lx0 = lsg_d * a_d - laref_d;
        for (;;) {
            float d1 = 0.f, d2 = 0.f;
            {
                const float xx = lx0 + Ls.al * ljv;
                if (lsg_d != 0.f && xx < 0.f) { d1 += lD_d * xx * ljv; d2 += lD_d * ljv * ljv; }
            }
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float xx = x[s][k] + Ls.al * jv[s][k];
                    if (xx < 0.f) { d1 += Dk[s] * xx * jv[s][k]; d2 += Dk[s] * jv[s][k] * jv[s][k]; }
                }
            d1 = warp_sum(d1) + q1 + Ls.al * q2; d2 = warp_sum(d2) + q2;
            if (!Ls.next(d1, d2)) break;
        }
        const float al = Ls.al;
        int changed = (lsg_d != 0.f) && ((lx0 < 0.f) != (lx0 + al * ljv < 0.f));
        a_d += al * s_d;
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float x1 = x[s][k] + al * jv[s][k];
                changed |= (Dk[s] != 0.f) && ((x[s][k] < 0.f) != (x1 < 0.f));
                x[s][k] = x1;
            }
        if (!__any_sync(kFull, changed) && fabsf(al - 1.f) < 1e-3f) { it++; break; }
    }
    iters_out = it;
    if (d < NV) a_out[d * a_stride] = a_d;
    __syncwarp();
}

}  // namespace gg
