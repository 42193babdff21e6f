// Coupled envs (a gripper pad may touch the cube or the plane): grid-wide list + warp-per-env drain, shared by the step kernels
// (included by gg_kernels.cu after StepArgs / reset_decision).
//
// Round 1 solved the coupled envs of a 32-env block with the three warps of that block (block-local teams): a block whose envs are all
// in contact kept its SM for ~100x longer than its neighbours and nothing balanced the load.  Now the kernel that finds a coupled env
// publishes its dynamics (gg::TeamWork, 0.9 KB) to the global scratch slab and appends the env to ONE list for the whole launch; when a
// block has finished its own envs, its warps pop envs from that list -- one WARP per env: narrow phase, Newton solve on all NU + 6 dofs
// (gg_coop.cuh team_solve), integration, observations, flags, counters -- until it is empty.  This is the compaction / persistent-warp
// scheme of stack_solve_kernel (gg_stack.cu) applied to the pick scene.
//
// List protocol.  No consumer ever waits for an entry that has not been appended (blocks that are not yet resident cannot be waited for:
// no deadlock on multi-wave grids), no compare-and-swap loops (1500 warps retrying a CAS on one word serialise: measured 10 ms per step),
// no "last block" bookkeeping:
//   producer  pos = atomicAdd(tail, k); clist[pos % cap ..] = envs; fence; cready[..] = epoch (release); atomicAdd(avail, k)
//   consumer  loop: if (avail <= 0) leave;  if (atomicSub(avail, 1) > 0) { t = atomicAdd(head, 1); wait cready[t % cap] == epoch; take clist[t % cap] }
//                   else { atomicAdd(avail, 1); }   and look at `avail` AGAIN before leaving
// No entry is lost: every append is followed by a drain pass of the appending block itself, and the consumer that performs the LAST
// look at `avail` of the launch sees it with no undo in flight, i.e. sees the true number of unclaimed entries and claims one if there
// is any.  tail / head grow monotonically over launches (ring indices modulo cap >= n_envs; the host zeroes them long before 2^31),
// cready carries the launch epoch, so nothing has to be re-armed at the end of a launch.
#pragma once

namespace cl {

enum { CTL_TAIL = 0, CTL_HEAD = 1, CTL_AVAIL = 2, CTL_COUPLED = kCtlCoupled };    // CTL_COUPLED: running count of coupled envs (both step kernels)

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// MAIN-side append of the warp's coupled lanes (cm = ballot of the coupled lanes; their TeamWork slots are already written)
__device__ __forceinline__ void list_append(const StepArgs& A, unsigned cm, bool coupled, int e0, int lane) {
    __threadfence();
    __syncwarp();
    unsigned base = 0;
    if (lane == 0) base = (unsigned)atomicAdd(&A.ctl[CTL_TAIL], __popc(cm));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (coupled) {
        const unsigned pos = (base + __popc(cm & ((1u << lane) - 1u))) % (unsigned)A.list_cap;
        A.clist[pos] = e0;
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.cready + pos), "r"(A.epoch) : "memory");
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); atomicAdd(&A.ctl[CTL_AVAIL], __popc(cm)); atomicAdd(&A.ctl[CTL_COUPLED], __popc(cm)); }
}

// claim one env of the list for this warp: env id, or -1 when the list is empty
__device__ __forceinline__ int list_pop(const StepArgs& A, int lane) {
    int e = -1;
    if (lane == 0) {
        while ((int)ld_relaxed(reinterpret_cast<const uint32_t*>(A.ctl + CTL_AVAIL)) > 0) {
            if (atomicSub(&A.ctl[CTL_AVAIL], 1) > 0) {
                const unsigned t = (unsigned)atomicAdd(&A.ctl[CTL_HEAD], 1) % (unsigned)A.list_cap;
                while (ld_relaxed(A.cready + t) != A.epoch) __nanosleep(32);
                e = (int)ld_relaxed(reinterpret_cast<const uint32_t*>(A.clist + t));
                break;
            }
            atomicAdd(&A.ctl[CTL_AVAIL], 1);              // nothing left for me: undo, then look again before leaving
        }
    }
    return __shfl_sync(0xffffffffu, e, 0);
}

// flags, counters and episode statistics of one env (CUBE role lanes and the lane-0 epilogue of a drained coupled env)
struct Tally { float v[8]; };      // steps, done, succ, return, length, nan, iterations, contacts
template <int NU>
__device__ __forceinline__ void commit_cube(const StepArgs& A, int e, gg::V3 cp, gg::Q4 cq, gg::V3 cv, gg::V3 cw, const float* qac, int steps, int epi,
                                            bool do_reset, bool active, Tally& T) {
    const size_t st = (size_t)A.stride;
    if (A.mode == 0 && !do_reset) { steps += 1; T.v[0] = 1.f; }
    const bool succ = cp.z > A.M.success_z;
    const float rew = succ ? 1.f : 0.f;
    const bool trunc = A.time_limit && steps >= A.M.horizon;
    if (A.reward) A.reward[e] = rew;
    if (A.terminated) A.terminated[e] = succ ? 1 : 0;
    if (A.truncated) A.truncated[e] = trunc ? 1 : 0;
    if (A.reward_host) A.reward_host[e] = rew;
    if (A.terminated_host) A.terminated_host[e] = succ ? 1 : 0;
    if (A.truncated_host) A.truncated_host[e] = trunc ? 1 : 0;
    if (A.mode != 2) {
        A.done[e] = (succ || trunc) ? 1 : 0;
        if (active || do_reset) {
            float* wp = A.state + e + (2 * NU) * st;
            wp[0] = cp.x; wp[st] = cp.y; wp[2 * st] = cp.z;
            wp[3 * st] = cq.w; wp[4 * st] = cq.x; wp[5 * st] = cq.y; wp[6 * st] = cq.z;
            wp[7 * st] = cv.x; wp[8 * st] = cv.y; wp[9 * st] = cv.z;
            wp[10 * st] = cw.x; wp[11 * st] = cw.y; wp[12 * st] = cw.z;
#pragma unroll
            for (int i = 0; i < 6; i++) wp[(13 + NU + i) * st] = qac[i];
        }
        A.step_count[e] = steps; A.episode[e] = epi;
        if (A.mode == 0 && (succ || trunc)) { T.v[1] = 1.f; T.v[2] = succ ? 1.f : 0.f; T.v[4] = (float)steps; }
        if (A.mode == 0) T.v[3] = rew;
    }
}


// ------------------------------------------------------------------------------------------------- coupled envs: drain
// One warp = one coupled env: warp-cooperative narrow phase + Newton solve on all NU + 6 dofs (gg_coop.cuh), then the env's
// integration, observations, flags and counters (every lane computes them redundantly from the broadcast solution, lane 0 stores).
template <int NA, int NF, bool FH>
__device__ __forceinline__ void drain_one(const StepArgs& A, int e, int lane, gg::TeamBuf<NA + NF + 6>& TB, gg::TeamWork<NA, NF>& TW, Tally& acc) {
    constexpr int NU = NA + NF, NV = NU + 6, AD = 7 + NF;
    constexpr int TWF = sizeof(gg::TeamWork<NA, NF>) / sizeof(float);
    static_assert(NV <= 32, "the solution is broadcast through TeamBuf::vec");
    float* twp = reinterpret_cast<float*>(&TW);
    float* avec = TB.vec;                                 // solution broadcast: NV floats (not used by team_solve)
    const float* src = A.team_scratch + (size_t)e * TWF;
#pragma unroll 1
    for (int i = lane; i < TWF; i += 32) twp[i] = __ldcg(src + i);
    __syncwarp();
    int its, nc;
    gg::team_solve<NA, NF, FH>(A.M, TW, TB, A.iterations, A.tol, lane, avec, 1, its, nc);
    __syncwarp();
    const uint32_t gid = (uint32_t)(A.first_env + e);
    bool do_reset = reset_decision(A, e, gid);
    int steps = A.step_count[e], epi = A.episode[e];
    if (do_reset) { epi += 1; steps = 0; }
    float q[NU], qd[NU], qa[NU], a[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) a[i] = avec[i];
#pragma unroll
    for (int i = 0; i < NU; i++) { q[i] = TW.q[i]; qd[i] = TW.qd[i]; qa[i] = TW.warm[i]; }
    gg::V3 cp = gg::mk(TW.cube[0], TW.cube[1], TW.cube[2]);
    gg::Q4 cq = gg::mkq(TW.cube[3], TW.cube[4], TW.cube[5], TW.cube[6]);
    gg::V3 cv = gg::mk(TW.cube[7], TW.cube[8], TW.cube[9]), cw = gg::mk(TW.cube[10], TW.cube[11], TW.cube[12]);
    float qac[6];
    gg::integrate_robot<NU>(A.M, q, qd, qa, a);            // coupled envs are active by construction (role_rnea)
    gg::integrate_cube(A.M, cp, cq, cv, cw, qac, a + NU);
    Tally T;
#pragma unroll
    for (int k = 0; k < 8; k++) T.v[k] = 0.f;
    {
        float acc = cp.x + cp.y + cp.z + cq.w + cq.x + cq.y + cq.z + cv.x + cv.y + cv.z + cw.x + cw.y + cw.z;
#pragma unroll
        for (int i = 0; i < NU; i++) acc += q[i] + qd[i];
        if (!isfinite(acc)) {                              // diverged env: re-spawn (same rule as the register path)
            float x, y;
            gg::spawn_draw(A.M, A.seed, gid, (uint32_t)epi, x, y);
            cp = gg::mk(x, y, A.M.spawn_z); cq = gg::ldq(A.M.spawn_quat);
            cv = gg::mk(0.f, 0.f, 0.f); cw = gg::mk(0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 6; i++) qac[i] = 0.f;
#pragma unroll
            for (int i = 0; i < NU; i++) { q[i] = A.M.home_q[i]; qd[i] = 0.f; qa[i] = 0.f; }
            epi += 1; steps = 0; T.v[5] = 1.f;
            do_reset = true;
        }
    }
    gg::V3 ep; gg::Q4 eq;
    gg::robot_eef<NA, NF>(A.M, q, ep, eq);
    if (lane == 0) {
        const size_t st = (size_t)A.stride;
        float oa[AD], oe[GG_ENV_STATE_DIM];
        oa[0] = ep.x; oa[1] = ep.y; oa[2] = ep.z; oa[3] = eq.w; oa[4] = eq.x; oa[5] = eq.y; oa[6] = eq.z;
#pragma unroll
        for (int f = 0; f < NF; f++) oa[7 + f] = q[NA + f];
        const gg::V3 d = ep - cp;
        oe[0] = cp.x; oe[1] = cp.y; oe[2] = cp.z; oe[3] = cq.w; oe[4] = cq.x; oe[5] = cq.y; oe[6] = cq.z;
        oe[7] = d.x; oe[8] = d.y; oe[9] = d.z; oe[10] = sqrtf(gg::dot(d, d));
#pragma unroll
        for (int i = 0; i < AD; i++) { A.obs_agent[(size_t)e * AD + i] = oa[i]; if (A.obs_agent_host) A.obs_agent_host[(size_t)e * AD + i] = oa[i]; }
#pragma unroll
        for (int i = 0; i < GG_ENV_STATE_DIM; i++) {
            A.obs_env[(size_t)e * GG_ENV_STATE_DIM + i] = oe[i];
            if (A.obs_env_host) A.obs_env_host[(size_t)e * GG_ENV_STATE_DIM + i] = oe[i];
        }
        float* wp = A.state + e;
#pragma unroll
        for (int i = 0; i < NU; i++) { wp[i * st] = q[i]; wp[(NU + i) * st] = qd[i]; wp[(2 * NU + 13 + i) * st] = qa[i]; }
        T.v[6] = (float)its; T.v[7] = (float)nc;
        commit_cube<NU>(A, e, cp, cq, cv, cw, qac, steps, epi, do_reset, true, T);
    }
    // episode statistics: accumulated per warp over the envs it drains (lane 0 holds them), flushed once by drain_all
#pragma unroll
    for (int k = 0; k < 8; k++) acc.v[k] += T.v[k];
    __syncwarp();
}

// the tail of every step kernel: this warp pops coupled envs until the list is empty
template <int NA, int NF, bool FH>
__device__ __forceinline__ void drain_all(const StepArgs& A, int lane, gg::TeamBuf<NA + NF + 6>& TB, gg::TeamWork<NA, NF>& TW) {
    Tally acc;
#pragma unroll
    for (int k = 0; k < 8; k++) acc.v[k] = 0.f;
    bool any = false;
    for (;;) {
        const int ce = list_pop(A, lane);
        if (ce < 0) break;
        drain_one<NA, NF, FH>(A, ce, lane, TB, TW, acc);
        any = true;
    }
    if (any && lane == 0 && A.stats != nullptr && A.mode == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) if (acc.v[k] != 0.f) atomicAdd(&A.stats[k], (double)acc.v[k]);
    }
}


}  // namespace cl
