// step_rs_kernel: the env.step kernel with ONE ROLE PER SM (included by gg_kernels.cu after StepArgs / reset_decision / store_slab).
//
// Why (profiles/r02_notes.md): the SM instruction cache holds 32 KB.  One env.step executes ~85 KB of SASS (MAIN 25 KB, RNEA 18 KB,
// CUBE 28 KB + prologues); with the three role warps of a 32-env block on the same SM every launch streams all of it through that
// cache and the SM issues ~1 instruction per cycle whatever its occupancy (measured: 36 us at 16384 envs, while each role ALONE runs
// in 14 - 22 us).  Here the grid is one persistent CTA per SM and a CTA runs a single role, so an SM only ever executes one role's
// code: it fits, and because the block -> SM placement of a plain launch is deterministic it stays cached from launch to launch.
//
//   CTA b: role = MAIN for b < n_main, RNEA for the next n_rnea CTAs, CUBE for the rest.  Role-CTA c owns the 32-env groups
//   g = c, c + n_role, c + 2 n_role, ...; its warp w takes every kWarps-th of them in increasing order (lane = env, as before).
//
// The roles exchange a handful of values per env.  They used to go through shared memory and named barriers inside a block; now
// they go through a per-group MAILBOX in global memory (L2 resident, [row][lane] floats, written with plain stores, read with
// ld.global.cg) guarded by four flags per group that carry the launch epoch (st.release.gpu / ld.acquire.gpu):
//   flag 0  C0   CUBE -> RNEA, MAIN : cube state after the masked reset (19 rows: pos3 quat4 linvel3 angvel3 warm6)
//   flag 1  FS   RNEA -> MAIN, CUBE : tau - bias (NU rows), coupled flag (1 row)
//   flag 2  ROB  MAIN -> CUBE       : robot non-finite flag, robot-block Newton iterations
//   flag 3  C1   CUBE -> MAIN       : final cube pose (7 rows), cube non-finite flag
// Every role processes its groups in increasing order and all CTAs are co-resident (cooperative launch), so the lowest unfinished
// group can always make progress: no deadlock.  MAIN writes agent_pos AND environment_state (it needs the end-effector position for
// the latter anyway), CUBE writes reward / flags / counters / statistics.
//
// Coupled envs (a gripper pad may touch the cube or the plane): MAIN publishes the env's dynamics (TeamWork) to the global scratch
// slab and appends the env to a grid-wide list; no role commits anything for it.  After its role loop every warp of every CTA turns
// into a solver team and drains that list ONE WARP PER ENV (gg_coop.cuh team_solve + integration + observations), i.e. the coupled
// envs of the whole batch are compacted and spread over all SMs instead of being solved by the three warps of their own block.
#pragma once

namespace rs {

using cl::Tally; using cl::commit_cube; using cl::drain_one; using cl::ld_relaxed;

#ifndef GG_RS_WARPS
#define GG_RS_WARPS 11
#endif
constexpr int kWarps = GG_RS_WARPS;                  // warps per CTA (one CTA per SM): 49 CTAs x 11 warps cover the 512 groups of 16384 envs in one pass
constexpr int kThreads = kWarps * 32;
constexpr int MB_C0 = 0, MB_FS = 19, MB_CPL = 28, MB_ROB = 29, MB_C1 = 31, MB_ROWS = 40;   // mailbox rows of one group
enum { F_C0 = 0, F_FS = 1, F_ROB = 2, F_C1 = 3 };
enum { CTL_TAIL = 0, CTL_HEAD = 1, CTL_MAIN_DONE = 2, CTL_ALL_DONE = 3 };

// Release / acquire WITHOUT L1 invalidation.  `__threadfence()` and `ld.acquire.gpu` both compile to ... + CCTL.IVALL on sm_100a, i.e. they
// drop the SM's whole L1 -- which holds the stack arrays (local memory) of every resident warp; with a dozen warps polling, the working
// warps lost their L1 every few hundred cycles and the first version of this kernel ran 2x slower than the kernel it replaces.
//   producer: every lane writes its LAST mailbox value with st.release.gpu (MEMBAR.ALL.GPU + strong store: the lane's earlier mailbox
//             stores are performed first), __syncwarp, lane 0 releases the flag;
//   consumer: lane 0 polls with ld.relaxed.gpu (a strong load, no fence), __syncwarp, all lanes read the mailbox with ld.global.cg
//             (L2, never L1: there is nothing stale to invalidate).
__device__ __forceinline__ void st_release_f32(float* p, float v) {
    asm volatile("st.release.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void post(uint32_t* f, uint32_t epoch, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) { return ld_relaxed(p); }   // see above: consumers read through L2 only
__device__ __forceinline__ void wait(const uint32_t* f, uint32_t epoch, int lane) {
#ifdef GG_RS_NOSYNC      // timing experiment (garbage results): no role ever waits for another one
    return;
#endif
    if (lane == 0) {
        while (ld_relaxed(f) != epoch) __nanosleep(32);
    }
    __syncwarp();
}

template <int NA, int NF> struct WarpScratch {      // per-warp shared memory: the largest of the four role views
    static constexpr int NU = NA + NF, AD = 7 + NF;
    static constexpr size_t main_b = (size_t)(gg::LinkStoreStrided<NA>::FLOATS + AD + GG_ENV_STATE_DIM) * 32 * sizeof(float);
    static constexpr size_t cube_b = (size_t)gg::CubeConStrided::FLOATS * 32 * sizeof(float);
    static constexpr size_t rnea_b = (size_t)NU * 32 * sizeof(float);
    static constexpr size_t team_b = sizeof(gg::TeamBuf<NU + 6>) + ((sizeof(gg::TeamWork<NA, NF>) + 15) / 16) * 16 + 64 * sizeof(float);
    static constexpr size_t m1 = main_b > cube_b ? main_b : cube_b;
    static constexpr size_t m2 = rnea_b > team_b ? rnea_b : team_b;
    static constexpr size_t bytes = ((m1 > m2 ? m1 : m2) + 127) / 128 * 128;
};

// -DGG_RS_TRACE: nanosecond timestamps of the first groups' milestones into the control block (developer instrumentation, tools/rs_trace.py)
#ifdef GG_RS_TRACE
__device__ __forceinline__ void trace(const StepArgs& A, int g, int role, int k, int lane) {
    if (lane == 0 && g < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* T = reinterpret_cast<unsigned long long*>(A.ctl + 64);
        T[(g * 3 + role) * 16 + k] = t;
        if (k == 0) T[(g * 3 + role) * 16 + 15] = smid;
    }
}
#define RS_TRACE(g, role, k) trace(A, g, role, k, lane)
#else
#define RS_TRACE(g, role, k)
#endif

struct Env {            // per-lane bookkeeping shared by the roles
    int e0, e; bool live, full, do_reset, active; uint32_t gid;
};
__device__ __forceinline__ Env env_of(const StepArgs& A, int g, int lane) {
    Env v;
    v.e0 = g * 32 + lane;
    v.live = v.e0 < A.n;
    v.e = v.live ? v.e0 : A.n - 1;                    // dead lanes shadow the last env (loads only)
    v.gid = (uint32_t)(A.first_env + v.e);
    v.full = g * 32 + 32 <= A.n;
    v.do_reset = reset_decision(A, v.e, v.gid);
    v.active = v.do_reset ? (A.M.settle_steps > 0) : (A.mode == 0);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------- RNEA
template <int NA, int NF, bool FH>
__device__ __forceinline__ void role_rnea(const StepArgs& A, int g, int lane, float* scratch) {
    constexpr int NU = NA + NF;
    RS_TRACE(g, 1, 0);
    const Env v = env_of(A, g, lane);
    float* mb = A.mbox + (size_t)g * MB_ROWS * 32 + lane;
    uint32_t* fl = A.mflag + (size_t)g * 4;
    const float* sp = A.state + v.e;
    const size_t st = (size_t)A.stride;
    // the group's (32, NU) slab of the action batch is fetched asynchronously into shared memory while the bias forces are computed
    // (HBM, or PCIe / C2C when the caller passes pinned host memory: gg_step_host, zero-copy)
    const bool staged = A.mode == 0 && A.stage_action != 0 && v.full;
    if (staged) {
        const float* src = A.action + (size_t)g * 32 * NU;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(scratch);
#pragma unroll
        for (int c = lane; c < 32 * NU / 4; c += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * c), "l"(src + 4 * c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    float q[NU], qd[NU], fs[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) {
        const float lq = sp[i * st], lqd = sp[(NU + i) * st];
        q[i] = v.do_reset ? A.M.home_q[i] : lq; qd[i] = v.do_reset ? 0.f : lqd;
    }
    ActionFetch<NU> tf;
    tf.home = A.M.home_q; tf.slab = scratch + lane * NU; tf.global = A.action + (size_t)v.e * NU;
    tf.staged = staged; tf.use_action = A.mode == 0 && !v.do_reset;
    gg::RobotDyn<NA, NF> D;                               // only the pad-body frames are filled (and kept)
    gg::RneaStoreLocal<NA> rs;
    gg::robot_rnea_t<NA, NF, FH>(A.M, A.C, q, qd, tf, fs, rs, &D);
    RS_TRACE(g, 1, 1);
#pragma unroll
    for (int i = 0; i < NU; i++) mb[(MB_FS + i) * 32] = fs[i];
    if (A.action_dev != nullptr && A.mode == 0) {         // device-resident copy of the consumed targets (host-I/O callers)
        if (staged) store_slab<32 * NU / 4>(scratch, A.action_dev + (size_t)g * 32 * NU, nullptr, lane);
        else if (v.live) {
#pragma unroll
            for (int i = 0; i < NU; i++) A.action_dev[(size_t)v.e * NU + i] = A.action[(size_t)v.e * NU + i];
        }
    }
    // pad broad phase against the cube state the CUBE role published (after its masked reset)
    wait(fl + F_C0, A.epoch, lane);
    RS_TRACE(g, 1, 2);
    const gg::V3 cp = gg::mk(__ldcg(mb + (MB_C0 + 0) * 32), __ldcg(mb + (MB_C0 + 1) * 32), __ldcg(mb + (MB_C0 + 2) * 32));
    const gg::Q4 cq = gg::mkq(__ldcg(mb + (MB_C0 + 3) * 32), __ldcg(mb + (MB_C0 + 4) * 32), __ldcg(mb + (MB_C0 + 5) * 32), __ldcg(mb + (MB_C0 + 6) * 32));
    const bool coupled = v.live && v.active && gg::pads_may_touch<NA, NF>(A.M, D, cp, cq);
    st_release_f32(mb + MB_CPL * 32, coupled ? 1.f : 0.f);
    post(fl + F_FS, A.epoch, lane);
    RS_TRACE(g, 1, 3);
    __syncwarp();                                          // the action slab in `scratch` is free for the next group from here on
}

// ---------------------------------------------------------------------------------------------------------------- MAIN
template <int NA, int NF, bool FH>
__device__ __forceinline__ void role_main(const StepArgs& A, int g, int lane, float* scratch) {
    constexpr int NU = NA + NF, AD = 7 + NF;
    constexpr int TWF = sizeof(gg::TeamWork<NA, NF>) / sizeof(float);
    RS_TRACE(g, 0, 0);
    Env v = env_of(A, g, lane);
    float* mb = A.mbox + (size_t)g * MB_ROWS * 32 + lane;
    uint32_t* fl = A.mflag + (size_t)g * 4;
    const float* sp = A.state + v.e;
    const size_t st = (size_t)A.stride;
    const bool run_sub = A.mode == 0 || (A.mode == 1 && A.M.settle_steps > 0);
    float* s_oa = scratch + gg::LinkStoreStrided<NA>::FLOATS * 32;      // agent_pos slab (32 x AD), then environment_state slab (32 x 11)
    float* s_oe = s_oa + 32 * AD;
    float q[NU], qd[NU], qa[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) {
        const float lq = sp[i * st], lqd = sp[(NU + i) * st], lqa = sp[(2 * NU + 13 + i) * st];
        q[i] = v.do_reset ? A.M.home_q[i] : lq; qd[i] = v.do_reset ? 0.f : lqd; qa[i] = v.do_reset ? 0.f : lqa;
    }
    bool coupled = false;
    int rob_iters = 0;
    if (run_sub) {
        gg::RobotDyn<NA, NF> D;
        {
            gg::LinkStoreStrided<NA> ls;                   // per-link temporaries: [element][lane] view of this warp's shared memory
            ls.b = scratch + lane;
            gg::robot_crba<NA, NF, FH>(A.M, A.C, q, D, ls);
        }
        RS_TRACE(g, 0, 7);
        float L[NU * (NU + 1) / 2], dinv[NU];
#pragma unroll
        for (int i = 0; i < NU * (NU + 1) / 2; i++) L[i] = D.Mt[i];
        gg::chol_packed<NU>(L, dinv);
        RS_TRACE(g, 0, 8);
        float lsg[NU], laref[NU], lD[NU];
        const int nlim = gg::joint_limit_rows<NU>(A.M, q, qd, lsg, laref, lD);
        RS_TRACE(g, 0, 1);
#ifdef GG_RS_STOP_A      // timing experiment: the MAIN role ends after CRBA + Cholesky + limit rows (its code then fits the instruction cache)
        mb[(MB_ROB + 1) * 32] = L[0] + L[44] + lsg[0] + laref[3] + lD[8] + dinv[2] + (float)nlim;
        return;
#endif
        wait(fl + F_FS, A.epoch, lane);
        RS_TRACE(g, 0, 2);
#pragma unroll
        for (int i = 0; i < NU; i++) D.fs[i] = __ldcg(mb + (MB_FS + i) * 32);
        coupled = __ldcg(mb + MB_CPL * 32) != 0.f;
#ifdef GG_RS_NOSYNC
        coupled = false;
#endif
        float as[NU], a[NU];
        gg::chol_solve_packed<NU>(L, dinv, D.fs, as);
        gg::StepInfo info;
        info.ncon = 0; info.iters = 0; info.grad = 0.f;
#pragma unroll
        for (int i = 0; i < NU; i++) a[i] = as[i];
        RS_TRACE(g, 0, 9);
        if (v.live && !coupled) gg::robot_block_solve<NU>(D.Mt, L, dinv, as, nlim, lsg, laref, lD, qa, a, A.iterations, A.tol, info);
        RS_TRACE(g, 0, 10);
        rob_iters = coupled ? 0 : info.iters;
        const unsigned cm = __ballot_sync(0xffffffffu, coupled);
        if (cm != 0u) {
            // publish the dynamics of the coupled envs and hand them to the grid-wide list (drained by every warp of every CTA)
            wait(fl + F_C0, A.epoch, lane);                // posted long ago (the broad phase already consumed it)
            if (coupled) {
                gg::TeamWork<NA, NF>& W = *reinterpret_cast<gg::TeamWork<NA, NF>*>(A.team_scratch + (size_t)v.e0 * TWF);
#pragma unroll
                for (int i = 0; i < NU * (NU + 1) / 2; i++) W.Mt[i] = D.Mt[i];
#pragma unroll
                for (int i = 0; i < NU; i++) {
                    W.as[i] = as[i]; W.qd[i] = qd[i]; W.warm[i] = qa[i]; W.lsg[i] = lsg[i]; W.laref[i] = laref[i]; W.lD[i] = lD[i];
                    W.q[i] = q[i];
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
                    W.fR[f][0] = D.fR[f].c0.x; W.fR[f][1] = D.fR[f].c0.y; W.fR[f][2] = D.fR[f].c0.z;
                    W.fR[f][3] = D.fR[f].c1.x; W.fR[f][4] = D.fR[f].c1.y; W.fR[f][5] = D.fR[f].c1.z;
                    W.fR[f][6] = D.fR[f].c2.x; W.fR[f][7] = D.fR[f].c2.y; W.fR[f][8] = D.fR[f].c2.z;
                    W.fp[f][0] = D.fp[f].x; W.fp[f][1] = D.fp[f].y; W.fp[f][2] = D.fp[f].z;
                }
                W.lR[0] = D.lR.c0.x; W.lR[1] = D.lR.c0.y; W.lR[2] = D.lR.c0.z;
                W.lR[3] = D.lR.c1.x; W.lR[4] = D.lR.c1.y; W.lR[5] = D.lR.c1.z;
                W.lR[6] = D.lR.c2.x; W.lR[7] = D.lR.c2.y; W.lR[8] = D.lR.c2.z;
                W.lp[0] = D.lp.x; W.lp[1] = D.lp.y; W.lp[2] = D.lp.z;
#pragma unroll
                for (int i = 0; i < 19; i++) W.cube[i] = __ldcg(mb + (MB_C0 + i) * 32);
            }
            __threadfence();
            __syncwarp();
            int base = 0;
            if (lane == 0) base = atomicAdd(&A.ctl[CTL_TAIL], __popc(cm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (coupled) {
                const int pos = base + __popc(cm & ((1u << lane) - 1u));
                A.clist[pos] = v.e0;
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.cready + pos), "r"(A.epoch) : "memory");
            }
        }
        if (v.active && !coupled) gg::integrate_robot<NU>(A.M, q, qd, qa, a);
    }
    bool bad_robot = false;
    {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NU; i++) acc += q[i] + qd[i];
        bad_robot = run_sub && v.active && v.live && !coupled && !isfinite(acc);
    }
    if (run_sub) {
        mb[MB_ROB * 32] = bad_robot ? 1.f : 0.f;
        st_release_f32(mb + (MB_ROB + 1) * 32, (float)rob_iters);
        post(fl + F_ROB, A.epoch, lane);
    }
    RS_TRACE(g, 0, 3);
    gg::V3 ep; gg::Q4 eq;
    gg::robot_eef<NA, NF>(A.M, q, ep, eq);
    RS_TRACE(g, 0, 4);
    wait(fl + F_C1, A.epoch, lane);
    RS_TRACE(g, 0, 5);
    const gg::V3 cp = gg::mk(__ldcg(mb + (MB_C1 + 0) * 32), __ldcg(mb + (MB_C1 + 1) * 32), __ldcg(mb + (MB_C1 + 2) * 32));
    const gg::Q4 cq = gg::mkq(__ldcg(mb + (MB_C1 + 3) * 32), __ldcg(mb + (MB_C1 + 4) * 32), __ldcg(mb + (MB_C1 + 5) * 32), __ldcg(mb + (MB_C1 + 6) * 32));
    const bool bad = bad_robot || (__ldcg(mb + (MB_C1 + 7) * 32) != 0.f);
    if (__any_sync(0xffffffffu, bad)) {                  // diverged env: re-spawn at the home pose (rare; the CUBE role re-spawns the cube)
        // NOTE: the home pose goes into a FRESH array selected element-wise; conditionally overwriting q[] in place
        // was miscompiled by nvcc 12.9 (stack copy of q[3..6] stored unconditionally), see profiles/r01_notes.md.
        float qh[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) qh[i] = bad ? A.M.home_q[i] : q[i];
        gg::V3 ep2; gg::Q4 eq2;
        gg::robot_eef<NA, NF>(A.M, qh, ep2, eq2);
#pragma unroll
        for (int i = 0; i < NU; i++) { q[i] = qh[i]; qd[i] = bad ? 0.f : qd[i]; qa[i] = bad ? 0.f : qa[i]; }
        ep = ep2; eq = eq2;
        v.do_reset = v.do_reset || bad;
    }
    const gg::V3 d = ep - cp;
    const float dn = sqrtf(gg::dot(d, d));
    const bool any_cpl = __any_sync(0xffffffffu, coupled);
    if (v.full && A.vec_out && !any_cpl) {
        // rows staged in shared memory, then one contiguous slab of 32 x AD (32 x 11) floats with 128-bit stores
        float* r = s_oa + lane * AD;
        r[0] = ep.x; r[1] = ep.y; r[2] = ep.z; r[3] = eq.w; r[4] = eq.x; r[5] = eq.y; r[6] = eq.z;
#pragma unroll
        for (int f = 0; f < NF; f++) r[7 + f] = q[NA + f];
        float* r2 = s_oe + lane * GG_ENV_STATE_DIM;
        r2[0] = cp.x; r2[1] = cp.y; r2[2] = cp.z; r2[3] = cq.w; r2[4] = cq.x; r2[5] = cq.y; r2[6] = cq.z;
        r2[7] = d.x; r2[8] = d.y; r2[9] = d.z; r2[10] = dn;
        __syncwarp();
        store_slab<32 * AD / 4>(s_oa, A.obs_agent + (size_t)g * 32 * AD, A.obs_agent_host ? A.obs_agent_host + (size_t)g * 32 * AD : nullptr, lane);
        store_slab<32 * GG_ENV_STATE_DIM / 4>(s_oe, A.obs_env + (size_t)g * 32 * GG_ENV_STATE_DIM,
                                              A.obs_env_host ? A.obs_env_host + (size_t)g * 32 * GG_ENV_STATE_DIM : nullptr, lane);
        __syncwarp();
    } else if (v.live && !coupled) {
        float oa[AD], oe[GG_ENV_STATE_DIM];
        oa[0] = ep.x; oa[1] = ep.y; oa[2] = ep.z; oa[3] = eq.w; oa[4] = eq.x; oa[5] = eq.y; oa[6] = eq.z;
#pragma unroll
        for (int f = 0; f < NF; f++) oa[7 + f] = q[NA + f];
        oe[0] = cp.x; oe[1] = cp.y; oe[2] = cp.z; oe[3] = cq.w; oe[4] = cq.x; oe[5] = cq.y; oe[6] = cq.z;
        oe[7] = d.x; oe[8] = d.y; oe[9] = d.z; oe[10] = dn;
#pragma unroll
        for (int i = 0; i < AD; i++) { A.obs_agent[(size_t)v.e * AD + i] = oa[i]; if (A.obs_agent_host) A.obs_agent_host[(size_t)v.e * AD + i] = oa[i]; }
#pragma unroll
        for (int i = 0; i < GG_ENV_STATE_DIM; i++) {
            A.obs_env[(size_t)v.e * GG_ENV_STATE_DIM + i] = oe[i];
            if (A.obs_env_host) A.obs_env_host[(size_t)v.e * GG_ENV_STATE_DIM + i] = oe[i];
        }
    }
    if (v.live && !coupled && A.mode != 2 && (v.active || v.do_reset)) {
        float* wp = A.state + v.e;
#pragma unroll
        for (int i = 0; i < NU; i++) { wp[i * st] = q[i]; wp[(NU + i) * st] = qd[i]; wp[(2 * NU + 13 + i) * st] = qa[i]; }
    }
    RS_TRACE(g, 0, 6);
}

// ---------------------------------------------------------------------------------------------------------------- CUBE
template <int NA, int NF, bool FH>
__device__ __forceinline__ void role_cube(const StepArgs& A, int g, int lane, float* scratch) {
    constexpr int NU = NA + NF;
    RS_TRACE(g, 2, 0);
    Env v = env_of(A, g, lane);
    float* mb = A.mbox + (size_t)g * MB_ROWS * 32 + lane;
    uint32_t* fl = A.mflag + (size_t)g * 4;
    const float* sp = A.state + v.e;
    const size_t st = (size_t)A.stride;
    const bool run_sub = A.mode == 0 || (A.mode == 1 && A.M.settle_steps > 0);
    Tally T;
#pragma unroll
    for (int k = 0; k < 8; k++) T.v[k] = 0.f;
    const float* cpn = sp + (2 * NU) * st;
    gg::V3 cp = gg::mk(cpn[0], cpn[st], cpn[2 * st]);
    gg::Q4 cq = gg::mkq(cpn[3 * st], cpn[4 * st], cpn[5 * st], cpn[6 * st]);
    gg::V3 cv = gg::mk(cpn[7 * st], cpn[8 * st], cpn[9 * st]);
    gg::V3 cw = gg::mk(cpn[10 * st], cpn[11 * st], cpn[12 * st]);
    float qac[6];
#pragma unroll
    for (int i = 0; i < 6; i++) qac[i] = sp[(2 * NU + 13 + NU + i) * st];
    int steps = A.step_count[v.e], epi = A.episode[v.e];
    if (v.do_reset) {
        float x, y;
        if (A.cube_xy != nullptr) { x = A.cube_xy[2 * v.e]; y = A.cube_xy[2 * v.e + 1]; }
        else gg::spawn_draw(A.M, A.seed, v.gid, (uint32_t)epi, x, y);
        cp = gg::mk(x, y, A.M.spawn_z); cq = gg::ldq(A.M.spawn_quat);
        cv = gg::mk(0.f, 0.f, 0.f); cw = gg::mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 6; i++) qac[i] = 0.f;
        epi += 1; steps = 0;
    }
    bool coupled = false;
    int ncon = 0, cube_iters = 0;
    if (run_sub) {
        mb[(MB_C0 + 0) * 32] = cp.x; mb[(MB_C0 + 1) * 32] = cp.y; mb[(MB_C0 + 2) * 32] = cp.z;
        mb[(MB_C0 + 3) * 32] = cq.w; mb[(MB_C0 + 4) * 32] = cq.x; mb[(MB_C0 + 5) * 32] = cq.y; mb[(MB_C0 + 6) * 32] = cq.z;
        mb[(MB_C0 + 7) * 32] = cv.x; mb[(MB_C0 + 8) * 32] = cv.y; mb[(MB_C0 + 9) * 32] = cv.z;
        mb[(MB_C0 + 10) * 32] = cw.x; mb[(MB_C0 + 11) * 32] = cw.y; mb[(MB_C0 + 12) * 32] = cw.z;
#pragma unroll
        for (int i = 0; i < 5; i++) mb[(MB_C0 + 13 + i) * 32] = qac[i];
        st_release_f32(mb + (MB_C0 + 18) * 32, qac[5]);
        post(fl + F_C0, A.epoch, lane);
        RS_TRACE(g, 2, 1);
        float Mc[6], as_c[6], qd_c[6], a_c[6];
        gg::cube_free_acc(A.M, cv, cw, Mc, as_c, qd_c);
        gg::StepInfo ic; ic.ncon = 0; ic.iters = 0; ic.grad = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) a_c[i] = as_c[i];
        if (v.live) {
            // speculative: the broad phase (end of the RNEA chain) has not reported yet; coupled envs drop this result
            gg::CubeConStrided cs;
            cs.b = scratch + lane;
            gg::cube_block_solve(A.M, cp, gg::qmat(cq), Mc, as_c, qd_c, qac, a_c, A.iterations, A.tol, cs, ic);
        }
        RS_TRACE(g, 2, 2);
        wait(fl + F_FS, A.epoch, lane);
        RS_TRACE(g, 2, 3);
        coupled = __ldcg(mb + MB_CPL * 32) != 0.f;
#ifdef GG_RS_NOSYNC
        coupled = false;
#endif
        ncon = ic.ncon; cube_iters = ic.iters;
        if (v.active && !coupled) gg::integrate_cube(A.M, cp, cq, cv, cw, qac, a_c);
    }
    bool bad_cube = false;
    {
        const float acc = cp.x + cp.y + cp.z + cq.w + cq.x + cq.y + cq.z + cv.x + cv.y + cv.z + cw.x + cw.y + cw.z;
        bad_cube = run_sub && v.active && v.live && !coupled && !isfinite(acc);
    }
    bool bad = bad_cube;
    int rob_iters = 0;
    if (run_sub) {
        RS_TRACE(g, 2, 4);
        wait(fl + F_ROB, A.epoch, lane);
        RS_TRACE(g, 2, 5);
        bad = bad || (__ldcg(mb + MB_ROB * 32) != 0.f);
        rob_iters = (int)__ldcg(mb + (MB_ROB + 1) * 32);
    }
    if (bad) {
        float x, y;
        gg::spawn_draw(A.M, A.seed, v.gid, (uint32_t)epi, x, y);
        cp = gg::mk(x, y, A.M.spawn_z); cq = gg::ldq(A.M.spawn_quat);
        cv = gg::mk(0.f, 0.f, 0.f); cw = gg::mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 6; i++) qac[i] = 0.f;
        epi += 1; steps = 0; T.v[5] = 1.f;
        v.do_reset = true;
    }
    mb[(MB_C1 + 0) * 32] = cp.x; mb[(MB_C1 + 1) * 32] = cp.y; mb[(MB_C1 + 2) * 32] = cp.z;
    mb[(MB_C1 + 3) * 32] = cq.w; mb[(MB_C1 + 4) * 32] = cq.x; mb[(MB_C1 + 5) * 32] = cq.y; mb[(MB_C1 + 6) * 32] = cq.z;
    st_release_f32(mb + (MB_C1 + 7) * 32, bad_cube ? 1.f : 0.f);
    post(fl + F_C1, A.epoch, lane);
    if (v.live && !coupled) {
        if (run_sub && v.active) { T.v[6] = (float)(cube_iters > rob_iters ? cube_iters : rob_iters); T.v[7] = (float)ncon; }
        commit_cube<NU>(A, v.e, cp, cq, cv, cw, qac, steps, epi, v.do_reset, v.active, T);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) T.v[k] = 0.f;
    }
    if (A.stats != nullptr && A.mode == 0) {
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float x = T.v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == k) mine = x;
        }
        if (lane < 8 && mine != 0.f) atomicAdd(&A.stats[lane], (double)mine);
    }
    RS_TRACE(g, 2, 6);
}

// Every warp ends here.  A warp always holds ONE ticket (fetch-and-add on the head: never fails, unlike a compare-and-swap loop, which
// serialised 1600 contending warps) and polls the ready word of ITS slot -- distinct addresses, no hot line.  A ticket beyond the
// final tail means the list is exhausted: the tail is final once every MAIN CTA has finished its role loop.
template <int NA, int NF, bool FH>
__device__ __forceinline__ void drain(const StepArgs& A, int lane, float* scratch) {
    const uint32_t* c_done = reinterpret_cast<const uint32_t*>(A.ctl + CTL_MAIN_DONE);
    const uint32_t* c_tail = reinterpret_cast<const uint32_t*>(A.ctl + CTL_TAIL);
    for (;;) {
        int e = -1;
        if (lane == 0) {
            const int t = atomicAdd(&A.ctl[CTL_HEAD], 1);
            unsigned ns = 64;
            for (;;) {
                if (ld_relaxed(A.cready + t) == A.epoch) { e = (int)ld_relaxed(reinterpret_cast<const uint32_t*>(A.clist + t)); break; }
                if ((int)ld_relaxed(c_done) == A.n_main && t >= (int)ld_relaxed(c_tail)) {
                    if (ld_relaxed(A.cready + t) == A.epoch) e = (int)ld_relaxed(reinterpret_cast<const uint32_t*>(A.clist + t));
                    break;
                }
                __nanosleep(ns);
                if (ns < 1024) ns *= 2;
            }
        }
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e < 0) break;
        {
            gg::TeamBuf<NA + NF + 6>& TBr = *reinterpret_cast<gg::TeamBuf<NA + NF + 6>*>(scratch);
            gg::TeamWork<NA, NF>& TWr = *reinterpret_cast<gg::TeamWork<NA, NF>*>(scratch + sizeof(gg::TeamBuf<NA + NF + 6>) / sizeof(float));
            Tally acc_;
#pragma unroll
            for (int k = 0; k < 8; k++) acc_.v[k] = 0.f;
            drain_one<NA, NF, FH>(A, e, lane, TBr, TWr, acc_);
            if (lane == 0 && A.stats != nullptr && A.mode == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) if (acc_.v[k] != 0.f) atomicAdd(&A.stats[k], (double)acc_.v[k]);
            }
        }
    }
}

template <int NA, int NF, bool FH>
__global__ void __launch_bounds__(kThreads, 1) step_rs_kernel(const __grid_constant__ StepArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* scratch = reinterpret_cast<float*>(smem_raw + (size_t)w * WarpScratch<NA, NF>::bytes);
    const int b = blockIdx.x;
    const int ngroups = (A.n + 31) / 32;
    const bool run_sub = A.mode == 0 || (A.mode == 1 && A.M.settle_steps > 0);     // grid-uniform
    int role, c, nr;
    if (b < A.n_main) { role = 0; c = b; nr = A.n_main; }
    else if (b < A.n_main + A.n_rnea) { role = 1; c = b - A.n_main; nr = A.n_rnea; }
    else { role = 2; c = b - A.n_main - A.n_rnea; nr = A.n_cube; }
    // this warp's groups, in increasing order: g = c + nr * (w + kWarps * k)
    if (role == 0) {
        for (int j = w; c + nr * j < ngroups; j += kWarps) role_main<NA, NF, FH>(A, c + nr * j, lane, scratch);
    } else if (role == 1) {
        if (run_sub)
            for (int j = w; c + nr * j < ngroups; j += kWarps) role_rnea<NA, NF, FH>(A, c + nr * j, lane, scratch);
    } else {
        for (int j = w; c + nr * j < ngroups; j += kWarps) role_cube<NA, NF, FH>(A, c + nr * j, lane, scratch);
    }
    __syncthreads();
    if (role == 0 && threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&A.ctl[CTL_MAIN_DONE], 1);
    }
    if (run_sub) drain<NA, NF, FH>(A, lane, scratch);
    // the last CTA to finish re-arms the list counters for the next launch (stream order makes that launch wait for this one)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&A.ctl[CTL_ALL_DONE], 1) == (int)gridDim.x - 1) {
            A.ctl[CTL_TAIL] = 0; A.ctl[CTL_HEAD] = 0; A.ctl[CTL_MAIN_DONE] = 0; A.ctl[CTL_ALL_DONE] = 0;
            __threadfence();
        }
    }
}

}  // namespace rs
