// CUDA kernels (sm_100a) and the C ABI of include/gg_b200.h.
//
// Mapping: one thread = one env; state is struct-of-arrays with the env index fastest so that every
// state load/store of a warp is one coalesced 128-byte line.  The whole env.step -- PD, FK, CRBA,
// Cholesky, RNEA, narrow phase, Newton solve, integration, obs / reward / flags, masked reset -- is
// ONE kernel launch; nothing is written back to HBM between those stages.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>

#include <atomic>
#include <new>

#include "gg_b200.h"
#include "gg_internal.h"
#include "gg_env.cuh"
#include "gg_coop.cuh"
#include "gg_ik.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return -1;
}
#define GG_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) return fail("CUDA error: %s", cudaGetErrorString(e_));  \
    } while (0)

struct StepArgs {
    GgModel M; gg::GgChain C;
    int n, stride;
    float* state; int32_t* step_count; int32_t* episode; uint8_t* done;
    const float* action; float* obs_agent; float* obs_env; float* reward;
    uint8_t* terminated; uint8_t* truncated;
    const uint8_t* reset_mask; const float* cube_xy;
    float* reward_host; uint8_t* terminated_host; uint8_t* truncated_host;   // nullable pinned-host mirrors (gg_step_host)
    float* obs_agent_host; float* obs_env_host;   // nullable pinned-host mirrors of the observations (gg_step_host)
    float* action_dev;        // nullable device copy of the consumed action batch (gg_step_host: PD targets persist on the device)
    int vec_out;              // 1: observation arrays (and their host mirrors) are 16-byte aligned -> full blocks write 128-bit slabs
    int stage_action;         // 1: action is 16-byte aligned -> full blocks stage their slab with cp.async
    int mode;                 // 0 = env.step, 1 = reset (+ settle) of masked envs, 2 = observe only
    int iterations; float tol;
    int auto_reset, time_limit; float reset_prob;
    uint64_t seed; int64_t first_env; uint32_t step_index;
    double* stats;
    float* team_scratch;      // n_envs x sizeof(TeamWork) / 4 floats: dynamics of coupled envs, published by their MAIN lane
    // role-per-SM kernel (gg_step_rs.cuh): per-group mailboxes + flags, launch epoch, coupled-env list, role split of the grid
    float* mbox; uint32_t* mflag; uint32_t epoch;
    int* ctl; int* clist; uint32_t* cready; int list_cap; int defer_drain;   // defer_drain: step_mono_kernel leaves the coupled list to drain_kernel
    int n_main, n_rnea, n_cube;
};

template <int NU> __device__ __forceinline__ void load_state(const float* __restrict__ st, int stride, int e, gg::EnvState<NU>& s) {
    const float* p = st + e;
#pragma unroll
    for (int i = 0; i < NU; i++) s.q[i] = p[(size_t)i * stride];
#pragma unroll
    for (int i = 0; i < NU; i++) s.qd[i] = p[(size_t)(NU + i) * stride];
    const float* c = p + (size_t)(2 * NU) * stride;
    s.cp = gg::mk(c[0], c[(size_t)stride], c[(size_t)2 * stride]);
    s.cq = gg::mkq(c[(size_t)3 * stride], c[(size_t)4 * stride], c[(size_t)5 * stride], c[(size_t)6 * stride]);
    s.cv = gg::mk(c[(size_t)7 * stride], c[(size_t)8 * stride], c[(size_t)9 * stride]);
    s.cw = gg::mk(c[(size_t)10 * stride], c[(size_t)11 * stride], c[(size_t)12 * stride]);
    const float* a = p + (size_t)(2 * NU + 13) * stride;
#pragma unroll
    for (int i = 0; i < NU + 6; i++) s.qacc[i] = a[(size_t)i * stride];
}

template <int NU> __device__ __forceinline__ void store_state(float* __restrict__ st, int stride, int e, const gg::EnvState<NU>& s) {
    float* p = st + e;
#pragma unroll
    for (int i = 0; i < NU; i++) p[(size_t)i * stride] = s.q[i];
#pragma unroll
    for (int i = 0; i < NU; i++) p[(size_t)(NU + i) * stride] = s.qd[i];
    float* c = p + (size_t)(2 * NU) * stride;
    c[0] = s.cp.x; c[(size_t)stride] = s.cp.y; c[(size_t)2 * stride] = s.cp.z;
    c[(size_t)3 * stride] = s.cq.w; c[(size_t)4 * stride] = s.cq.x; c[(size_t)5 * stride] = s.cq.y; c[(size_t)6 * stride] = s.cq.z;
    c[(size_t)7 * stride] = s.cv.x; c[(size_t)8 * stride] = s.cv.y; c[(size_t)9 * stride] = s.cv.z;
    c[(size_t)10 * stride] = s.cw.x; c[(size_t)11 * stride] = s.cw.y; c[(size_t)12 * stride] = s.cw.z;
    float* a = p + (size_t)(2 * NU + 13) * stride;
#pragma unroll
    for (int i = 0; i < NU + 6; i++) a[(size_t)i * stride] = s.qacc[i];
}

// ------------------------------------------------------------------------------------------------
// step_kernel: ROLE-SPECIALISED warps.  A block owns 32 consecutive envs and runs three warps on them;
// lane = env, warp = role:
//   warp 0  MAIN : kinematics + CRBA + Cholesky, smooth acceleration, joint-limit rows, robot-block Newton solve,
//                  robot integration, end-effector pose, agent_pos
//   warp 1  RNEA : velocities / accelerations / link forces of the same chain, PD torque -> fs = tau - bias; its chain walk
//                  yields the frames of the pad-carrying bodies, so it also runs the pad broad phase (coupled flags)
//   warp 2  CUBE : masked reset / spawn, cube-plane contacts + cube-block Newton solve, cube integration, environment_state,
//                  reward / terminated / truncated / done, counters, episode statistics
// The three chains are independent until the points marked by the named barriers below, so the latency of one env.step
// is the longest role instead of their sum, and 16384 envs put 2.6 warps on every scheduler instead of 0.9.
// Exchange goes through shared memory ([field][lane], conflict free):
//   barrier 8 (96 threads): CUBE arrives, RNEA and MAIN wait      -- cube state (after a masked reset) for the broad phase
//   barrier 1 (96 threads): RNEA arrives, MAIN and CUBE wait      -- fs, coupled flags
//   barrier 2 (64 threads): MAIN arrives, CUBE waits              -- solver statistics of the robot block
//   barrier 3 (64 threads): MAIN and CUBE wait                    -- end-effector position, non-finite flags
//   barrier 6 (64 threads): MAIN and RNEA wait                    -- integrated joint positions for the split end-effector chain
//   barrier 7 (64 threads): RNEA arrives, MAIN waits              -- tail of the end-effector chain
// The CUBE warp solves its cube block SPECULATIVELY (it does not wait for the broad phase, which sits at the end of the RNEA
// chain); the result is dropped for the envs that turn out to be coupled.
//
// Envs whose pads may touch the cube or the plane ("coupled") leave the
// thread-per-env mapping: after barrier 1 the three warps of the block become three solver teams and work through the
// block's coupled envs ONE WARP PER ENV (gg_coop.cuh; 10 KB of shared memory per team).  The MAIN lane of a coupled env
// publishes its dynamics to the env's slot of a global scratch slab (0.9 KB, L2 resident); after barrier 4 the teams pop
// envs from a block-local queue (team_queue) until it is empty, barrier 5.  Blocks without coupled envs never execute
// that code.  Tried and measured: team buffers in a global scratch slice -- 1.5x slower teams; a device-launched second kernel
// -- 2.6x slower (profiles/r01_notes.md); ONE grid-wide list of coupled envs drained by every warp of the launch (gg_drain.cuh, the
// scheme of the thread-per-env kernel below) -- 19.6 M instead of 32 M env-steps/s on the scripted grasp: 1500 warps at unrelated
// points of the 155 KB solver stream their instructions separately, while the block-local teams of a launch start together and share
// most fetches (SM IPC 0.69 vs 0.95, stall_no_instruction 61 % of the solver's samples; profiles/r02_notes.md).
//
// Outputs: agent_pos (32 x (7 + NF) floats) and environment_state (32 x 11 floats) of a full block are contiguous slabs of the
// env-major output arrays; the lanes stage their rows in shared memory and the warp writes the slab with 128-bit stores (and,
// for gg_step_host callers that ask for them, the same slabs into pinned host memory).  The action slab is staged the other
// way with 16-byte cp.async and, on request, copied to a device-resident target buffer (PD targets persist across steps).
// ------------------------------------------------------------------------------------------------
constexpr int kBlockThreads = 96;
constexpr int kCtlCoupled = 3;         // StepArgs::ctl slot: running count of coupled envs (both step kernels; read back lagged by launch_step)

#ifndef GG_MINBLOCKS
#define GG_MINBLOCKS 4
#endif
// -DGG_ONLY_ROLE=r (timing experiment, results are garbage): only role r runs and the named barriers are no-ops -- measures one
// role's chain when an SM executes nothing but that role's code (instruction-cache footprint experiment, profiles/r02_notes.md)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    __syncwarp();
#ifndef GG_ONLY_ROLE
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    __threadfence_block();
    __syncwarp();
#ifndef GG_ONLY_ROLE
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

template <int NU, int AD> struct Xchg {
    alignas(16) float act[32 * NU];   // this block's (32, NU) slab of the action batch, staged by the RNEA warp with cp.async
    alignas(16) float oa[32 * AD];    // agent_pos slab of the block (MAIN), written out with 128-bit stores
    alignas(16) float oe[32 * GG_ENV_STATE_DIM];   // environment_state slab (CUBE)
    float fs[NU][32];         // RNEA -> MAIN
    float cube[19][32];       // CUBE -> RNEA / MAIN: cp3 cq4 cv3 cw3 qacc6 (the state a coupled solve starts from)
    float aout[NU + 6][32];   // solver teams -> MAIN / CUBE: acceleration of coupled envs
    float ep[3][32];          // MAIN -> CUBE: end-effector position after the step
    float tail[7][32];        // RNEA -> MAIN: end-effector pose relative to arm link KS - 1 after the step (split chain)
    int coupled[32];          // RNEA -> MAIN / CUBE
    int iters[32], ncon[32];  // MAIN / teams -> CUBE (statistics)
    int bad[2][32];           // non-finite flags: [0] robot (MAIN), [1] cube (CUBE)
    int next;                 // head of the block's queue of coupled envs (solver teams pop with atomicAdd)
    alignas(8) unsigned long long mbar;   // transaction barrier of the bulk copy that stages `act`
};

// shared memory of one group of 32 envs (three role warps); a block holds SUB of them (dynamic shared memory)
template <int NA, int NF> struct StepSmem {
    // until the first team round the team buffers double as [element][lane] scratch of the rolled loops of MAIN (per-link S, I) and CUBE
    // (per-contact rows); `pad` makes the region starting at TB large enough for both views
    static constexpr size_t kScratch = (size_t)(gg::LinkStoreStrided<NA>::FLOATS + gg::CubeConStrided::FLOATS) * 32 * sizeof(float);
    static constexpr size_t kTB = 3 * sizeof(gg::TeamBuf<NA + NF + 6>);
    Xchg<NA + NF, 7 + NF> X;
    gg::TeamWork<NA, NF> TW[3];
    gg::TeamBuf<NA + NF + 6> TB[3];
    float pad[kScratch > kTB ? (kScratch - kTB + 3) / 4 : 1];
};

// PD targets of one env for robot_rnea_t: the staged slab in shared memory (after the cp.async group has landed), the action
// row in global memory (partial tail block / unaligned batch), or the home pose (envs being reset, settle steps).
template <int NU> struct ActionFetch {
    const float* home; const float* slab; const float* global;
    bool staged, use_action;
    unsigned mbar = 0;                                  // shared-memory address of the transaction barrier (bulk copy), 0: cp.async groups
    __device__ __forceinline__ void ready() {
        if (staged) {                                   // block-uniform
            if (mbar != 0u) {
                unsigned done = 0;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
                } while (!done);
            } else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
    }
    __device__ __forceinline__ float operator()(int d) const {
        const float a = staged ? slab[d] : (use_action ? global[d] : 0.f);
        return use_action ? a : home[d];
    }
};

__device__ __forceinline__ bool reset_decision(const StepArgs& A, int e, uint32_t gid) {
    bool r = false;
    if (A.mode == 0) {
        r = (A.reset_mask != nullptr && A.reset_mask[e] != 0) || (A.auto_reset && A.done[e] != 0);
        if (A.reset_prob > 0.f) {
            uint32_t o[4];
            gg::philox4x32_10(gid, A.step_index, 0x52535421u, 0u, (uint32_t)A.seed, (uint32_t)(A.seed >> 32), o);
            r = r || gg::u01(o[0]) < A.reset_prob;
        }
    } else if (A.mode == 1) {
        r = A.reset_mask == nullptr || A.reset_mask[e] != 0;
    }
    return r;
}

// a warp writes its block's contiguous slab (NF4 float4's staged in shared memory) to one or two destinations
template <int NF4> __device__ __forceinline__ void store_slab(const float* smem, float* dst, float* dst_host, int lane) {
    const float4* s4 = reinterpret_cast<const float4*>(smem);
#pragma unroll
    for (int c = lane; c < NF4; c += 32) {
        const float4 v = s4[c];
        reinterpret_cast<float4*>(dst)[c] = v;
        if (dst_host) reinterpret_cast<float4*>(dst_host)[c] = v;
    }
}

// Solver-team work queue: the three warps of the block pop the block's coupled envs one at a time (X.next), pull the env's
// published dynamics from the global scratch slab into their shared-memory TeamWork (7 coalesced 128-byte lines) and solve it
// warp-cooperatively.  No barrier between envs: a team that drew a cheap env (few contacts, warm start already optimal) moves
// straight on to the next one instead of waiting for the slowest of three (the lock-step rounds this replaces spent 27 % of
// their stall samples in bar.sync, profiles/r01_v7_grasp_*).
template <int NA, int NF, bool FH, class XT>
__device__ __forceinline__ void team_queue(const StepArgs& A, gg::TeamWork<NA, NF>* TW, gg::TeamBuf<NA + NF + 6>* TB, XT& X, unsigned cm,
                                           int ncpl, int team, int lane, int blk) {
    constexpr int TWF = sizeof(gg::TeamWork<NA, NF>) / sizeof(float);
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(&X.next, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= ncpl) break;
        const int el = __fns(cm, 0, k + 1);
        const float* src = A.team_scratch + (size_t)(blk * 32 + el) * TWF;
        float* dst = reinterpret_cast<float*>(&TW[team]);
#pragma unroll 1
        for (int i = lane; i < TWF; i += 32) dst[i] = __ldcg(src + i);
        __syncwarp();
        int its, nc;
        gg::team_solve<NA, NF, FH>(A.M, TW[team], TB[team], A.iterations, A.tol, lane, &X.aout[0][el], 32, its, nc);
        if (lane == 0) { X.iters[el] = its; X.ncon[el] = nc; }
        __syncwarp();
    }
    __threadfence_block();
}

// MB = resident blocks per SM the register allocation aims at: 4 (168 registers) for launches that need them (16384 envs = 3.5 blocks per SM),
// 2 (246 registers: the packed Cholesky factor and the limit rows of the MAIN role stop spilling) for launches of at most two blocks per SM
// (4096 envs: 31.7 -> 27.9 us, 8192 envs: 33.7 -> 29.0 us; profiles/r02_notes.md)
template <int NA, int NF, bool FH, int SUB, int MB>
__global__ void __launch_bounds__(kBlockThreads * SUB, MB / SUB) step_kernel(const __grid_constant__ StepArgs A) {
    constexpr int NU = NA + NF, NV = NU + 6, AD = 7 + NF;
    constexpr int KS = (NA - 1) / 2;      // split point of the end-effector chain (links [0, KS) MAIN, [KS, NA) RNEA)
    typedef Xchg<NU, AD> XT;
    typedef StepSmem<NA, NF> SM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // until the first team round the team buffers double as scratch for the rolled loops of MAIN (per-link S, I) and CUBE
    // (per-contact J, aref, D, u, w)
    static_assert(sizeof(SM::TB) + sizeof(SM::pad) >= SM::kScratch && offsetof(SM, pad) == offsetof(SM, TB) + sizeof(SM::TB), "team buffers too small for the scratch views");
    // SUB groups of 32 envs per block: warp w = role * SUB + sub, so the SUB warps of one role sit on different schedulers and meet
    // at the same (SUB times wider) named barriers -- they run the role's instruction stream in step (see the header comment)
    const int lane = threadIdx.x & 31, role = (threadIdx.x >> 5) / SUB, sub = (threadIdx.x >> 5) % SUB;
    const int blk = blockIdx.x * SUB + sub;               // index of this warp's group of 32 envs
    const int pbar = 9 + sub;                             // named barrier private to the three warps of this group
    SM& S = reinterpret_cast<SM*>(smem_raw)[sub];
    XT& X = S.X;
    gg::TeamWork<NA, NF>* TW = S.TW;
    gg::TeamBuf<NV>* TB = S.TB;
#ifdef GG_ONLY_ROLE
    if (role != GG_ONLY_ROLE) return;
    X.coupled[lane] = 0;
    __syncwarp();
#endif
    if (role == 2 && lane == 0) X.next = 0;                    // CUBE lane 0: ordered before every pop by barrier 8 (arrive) and barrier 4
    const int e0 = blk * 32 + lane;
    const bool live = e0 < A.n;
    const int e = live ? e0 : A.n - 1;                    // dead lanes shadow the last env (loads only)
    const uint32_t gid = (uint32_t)(A.first_env + e);
    const bool run_sub = A.mode == 0 || (A.mode == 1 && A.M.settle_steps > 0);     // block-uniform
    const bool full = blk * 32 + 32 <= A.n;                                 // block-uniform: no dead lanes
    bool do_reset = reset_decision(A, e, gid);
    const bool active = do_reset ? (A.M.settle_steps > 0) : (A.mode == 0);         // this env runs a substep
    const float* sp = A.state + e;
    const size_t st = (size_t)A.stride;

    if (role == 1) {
        // ------------------------------------------------------------------------------ RNEA
        if (!run_sub) return;
        // The action slab of this block (32 x NU floats, contiguous) is fetched asynchronously into shared memory while the
        // bias forces are computed: the targets are only needed by the PD law at the very end of this role, so the fetch
        // latency -- HBM, or PCIe / C2C when the caller passes pinned host memory (gg_step_host, zero-copy) -- is hidden.
#ifndef GG_NO_STAGE
        const bool staged = A.mode == 0 && A.stage_action != 0 && full;      // block-uniform
#else
        const bool staged = false;
#endif
        unsigned mbar_addr = 0;
        if (staged) {
            const float* src = A.action + (size_t)blk * 32 * NU;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(X.act);
#ifndef GG_ACTION_CP_ASYNC
            // ONE bulk copy of the block's 32 x NU slab (1152 B for the Panda) with a transaction barrier: when the batch lives in pinned
            // host memory the read crosses PCIe as a few large requests instead of 72 16-byte ones
            constexpr unsigned kBytes = 32u * NU * sizeof(float);
            mbar_addr = (unsigned)__cvta_generic_to_shared(&X.mbar);
            if (lane == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"(kBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(src), "r"(kBytes), "r"(mbar_addr) : "memory");
            }
            __syncwarp();
#else
#pragma unroll
            for (int c = lane; c < 32 * NU / 4; c += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * c), "l"(src + 4 * c) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
#endif
        }
        float q[NU], qd[NU], fs[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) {
            const float lq = sp[i * st], lqd = sp[(NU + i) * st];
            q[i] = do_reset ? A.M.home_q[i] : lq; qd[i] = do_reset ? 0.f : lqd;
        }
        ActionFetch<NU> tf;
        tf.home = A.M.home_q; tf.slab = X.act + lane * NU; tf.global = A.action + (size_t)e * NU;
        tf.staged = staged; tf.use_action = A.mode == 0 && !do_reset; tf.mbar = mbar_addr;
        gg::RobotDyn<NA, NF> D;                               // only the pad-body frames are filled (and kept)
        gg::RneaStoreLocal<NA> rs;
        gg::robot_rnea_t<NA, NF, FH>(A.M, A.C, q, qd, tf, fs, rs, &D);
#pragma unroll
        for (int i = 0; i < NU; i++) X.fs[i][lane] = fs[i];
        if (A.action_dev != nullptr && A.mode == 0) {         // device-resident copy of the consumed targets (host-I/O callers)
            if (staged) store_slab<32 * NU / 4>(X.act, A.action_dev + (size_t)blk * 32 * NU, nullptr, lane);
            else if (live) {
#pragma unroll
                for (int i = 0; i < NU; i++) A.action_dev[(size_t)e * NU + i] = A.action[(size_t)e * NU + i];
            }
        }
        // pad broad phase against the cube state the CUBE warp published (after its masked reset)
        bar_sync(8, 96 * SUB);
        const gg::V3 cp = gg::mk(X.cube[0][lane], X.cube[1][lane], X.cube[2][lane]);
        const gg::Q4 cq = gg::mkq(X.cube[3][lane], X.cube[4][lane], X.cube[5][lane], X.cube[6][lane]);
#ifdef GG_ONLY_ROLE
        const bool coupled = live && gg::pads_may_touch<NA, NF>(A.M, D, cp, cq) && cp.x > 1e30f;
#else
        const bool coupled = live && gg::pads_may_touch<NA, NF>(A.M, D, cp, cq);
#endif
        X.coupled[lane] = coupled ? 1 : 0;
        bar_arrive(1, 96 * SUB);
        const unsigned cm = __ballot_sync(0xffffffffu, coupled);
        const int ncpl = __popc(cm);
        if (ncpl > 0) {                                   // solver team 1 (see MAIN for the protocol)
            bar_sync(pbar, 96);
            team_queue<NA, NF, FH>(A, TW, TB, X, cm, ncpl, 1, lane, blk);
            bar_sync(pbar, 96);
        }
#ifndef GG_NO_SPLIT_EEF
        {
            // post-step refresh of the end-effector pose, second half of the chain (MAIN walks the first half meanwhile)
            bar_sync(6, 64 * SUB);
            float qt[NA - KS];
#pragma unroll
            for (int i = 0; i < NA - KS; i++) qt[i] = X.fs[KS + i][lane];
            gg::V3 tp; gg::Q4 tq;
            gg::robot_chain_tail<NA>(A.M, qt, KS, tp, tq);
            X.tail[0][lane] = tp.x; X.tail[1][lane] = tp.y; X.tail[2][lane] = tp.z;
            X.tail[3][lane] = tq.w; X.tail[4][lane] = tq.x; X.tail[5][lane] = tq.y; X.tail[6][lane] = tq.z;
            bar_arrive(7, 64 * SUB);
        }
#endif
        return;
    }

    if (role == 0) {
        // ------------------------------------------------------------------------------ MAIN
        float q[NU], qd[NU], qa[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) {
            const float lq = sp[i * st], lqd = sp[(NU + i) * st], lqa = sp[(2 * NU + 13 + i) * st];
            q[i] = do_reset ? A.M.home_q[i] : lq; qd[i] = do_reset ? 0.f : lqd; qa[i] = do_reset ? 0.f : lqa;
        }
        bool commit = live;
        if (run_sub) {
            gg::RobotDyn<NA, NF> D;
            {
                // per-link temporaries in the (still idle) team buffers: [element][lane] view of shared memory
                gg::LinkStoreStrided<NA> ls;
                ls.b = reinterpret_cast<float*>(TB) + lane;
                gg::robot_crba<NA, NF, FH>(A.M, A.C, q, D, ls);
            }
            float L[NU * (NU + 1) / 2], dinv[NU];
#pragma unroll
            for (int i = 0; i < NU * (NU + 1) / 2; i++) L[i] = D.Mt[i];
            gg::chol_packed<NU>(L, dinv);
            float lsg[NU], laref[NU], lD[NU];
            const int nlim = gg::joint_limit_rows<NU>(A.M, q, qd, lsg, laref, lD);
            bar_sync(8, 96 * SUB);
            bar_sync(1, 96 * SUB);
#pragma unroll
            for (int i = 0; i < NU; i++) D.fs[i] = X.fs[i][lane];
            float as[NU], a[NU];
            gg::chol_solve_packed<NU>(L, dinv, D.fs, as);
            gg::StepInfo info;
            info.ncon = 0; info.iters = 0; info.grad = 0.f;
            const bool coupled = X.coupled[lane] != 0;
#pragma unroll
            for (int i = 0; i < NU; i++) a[i] = as[i];
            if (commit && !coupled) gg::robot_block_solve<NU>(D.Mt, L, dinv, as, nlim, lsg, laref, lD, qa, a, A.iterations, A.tol, info);
            {
                const unsigned cm = __ballot_sync(0xffffffffu, coupled);
                const int ncpl = __popc(cm);
                if (ncpl > 0) {
                    if (lane == 0) atomicAdd(&A.ctl[kCtlCoupled], ncpl);     // running count for the host's kernel choice (launch_step)
                    // protocol: the MAIN lane of every coupled env publishes its dynamics to the env's slot of the global scratch
                    // slab (L2 resident), barrier 4 (the team buffers are free of MAIN's / CUBE's scratch views from here on),
                    // the three warps drain the block's queue of coupled envs, barrier 5 (accelerations are in X.aout)
                    if (coupled) {
                        constexpr int TWF = sizeof(gg::TeamWork<NA, NF>) / sizeof(float);
                        gg::TeamWork<NA, NF>& W = *reinterpret_cast<gg::TeamWork<NA, NF>*>(A.team_scratch + (size_t)e0 * TWF);
#pragma unroll
                        for (int i = 0; i < NU * (NU + 1) / 2; i++) W.Mt[i] = D.Mt[i];
#pragma unroll
                        for (int i = 0; i < NU; i++) {
                            W.as[i] = as[i]; W.qd[i] = qd[i]; W.warm[i] = qa[i]; W.lsg[i] = lsg[i]; W.laref[i] = laref[i]; W.lD[i] = lD[i];
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
                        for (int i = 0; i < 19; i++) W.cube[i] = X.cube[i][lane];
                    }
                    __threadfence_block();
                    bar_sync(pbar, 96);
                    team_queue<NA, NF, FH>(A, TW, TB, X, cm, ncpl, 0, lane, blk);
                    bar_sync(pbar, 96);
                }
                if (coupled) {
#pragma unroll
                    for (int i = 0; i < NU; i++) a[i] = X.aout[i][lane];
                }
            }
            if (!coupled) X.iters[lane] = info.iters;
            bar_arrive(2, 64 * SUB);
            if (active) gg::integrate_robot<NU>(A.M, q, qd, qa, a);
        }
        gg::V3 ep; gg::Q4 eq;
#ifndef GG_NO_SPLIT_EEF
        if (run_sub) {
            // split chain: the (by now idle) RNEA warp walks links [KS, NA) + the end-effector offset, this warp links [0, KS)
#pragma unroll
            for (int i = KS; i < NA; i++) X.fs[i][lane] = q[i];
            __threadfence_block();
            bar_sync(6, 64 * SUB);
            gg::Q4 ql; gg::V3 pl; gg::M3 Rl;
            gg::robot_chain_head<NA>(A.M, q, KS, ql, pl, Rl);
            bar_sync(7, 64 * SUB);
            gg::compose_eef(ql, pl, Rl, gg::mk(X.tail[0][lane], X.tail[1][lane], X.tail[2][lane]),
                            gg::mkq(X.tail[3][lane], X.tail[4][lane], X.tail[5][lane], X.tail[6][lane]), ep, eq);
        } else
#endif
        gg::robot_eef<NA, NF>(A.M, q, ep, eq);
        {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < NU; i++) acc += q[i] + qd[i];
            X.bad[0][lane] = (run_sub && active && commit && !isfinite(acc)) ? 1 : 0;
        }
        X.ep[0][lane] = ep.x; X.ep[1][lane] = ep.y; X.ep[2][lane] = ep.z;
        __threadfence_block();
        bar_sync(3, 64 * SUB);
        const bool bad = (X.bad[0][lane] | X.bad[1][lane]) != 0;
        if (__any_sync(0xffffffffu, bad)) {              // diverged env: re-spawn (rare; both roles take this branch together)
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
            X.ep[0][lane] = ep.x; X.ep[1][lane] = ep.y; X.ep[2][lane] = ep.z;
            do_reset = do_reset || bad;
            __threadfence_block();
            bar_sync(pbar, 64);
        }
        if (full && A.vec_out) {
            // agent_pos rows staged in shared memory (row stride AD words: conflict free for the Panda's 9, a 4-way conflict on
            // eight stores for the SO-101's 8), then one contiguous slab of 32 x AD floats with 128-bit stores
            float* r = X.oa + lane * AD;
            r[0] = ep.x; r[1] = ep.y; r[2] = ep.z; r[3] = eq.w; r[4] = eq.x; r[5] = eq.y; r[6] = eq.z;
#pragma unroll
            for (int f = 0; f < NF; f++) r[7 + f] = q[NA + f];
            __syncwarp();
            store_slab<32 * AD / 4>(X.oa, A.obs_agent + (size_t)blk * 32 * AD,
                                    A.obs_agent_host ? A.obs_agent_host + (size_t)blk * 32 * AD : nullptr, lane);
        } else if (commit) {
            float* oa = A.obs_agent + (size_t)e * AD;
            oa[0] = ep.x; oa[1] = ep.y; oa[2] = ep.z; oa[3] = eq.w; oa[4] = eq.x; oa[5] = eq.y; oa[6] = eq.z;
#pragma unroll
            for (int f = 0; f < NF; f++) oa[7 + f] = q[NA + f];
            if (A.obs_agent_host) {
                float* oh = A.obs_agent_host + (size_t)e * AD;
                oh[0] = ep.x; oh[1] = ep.y; oh[2] = ep.z; oh[3] = eq.w; oh[4] = eq.x; oh[5] = eq.y; oh[6] = eq.z;
#pragma unroll
                for (int f = 0; f < NF; f++) oh[7 + f] = q[NA + f];
            }
        }
        if (commit && A.mode != 2 && (active || do_reset)) {
            float* wp = A.state + e;
#pragma unroll
            for (int i = 0; i < NU; i++) { wp[i * st] = q[i]; wp[(NU + i) * st] = qd[i]; wp[(2 * NU + 13 + i) * st] = qa[i]; }
        }
        return;
    }

    // ---------------------------------------------------------------------------------- CUBE
    float st_steps = 0.f, st_done = 0.f, st_succ = 0.f, st_ret = 0.f, st_len = 0.f, st_nan = 0.f, st_iter = 0.f, st_con = 0.f;
    {
        const float* cpn = sp + (2 * NU) * st;
        gg::V3 cp = gg::mk(cpn[0], cpn[st], cpn[2 * st]);
        gg::Q4 cq = gg::mkq(cpn[3 * st], cpn[4 * st], cpn[5 * st], cpn[6 * st]);
        gg::V3 cv = gg::mk(cpn[7 * st], cpn[8 * st], cpn[9 * st]);
        gg::V3 cw = gg::mk(cpn[10 * st], cpn[11 * st], cpn[12 * st]);
        float qac[6];
#pragma unroll
        for (int i = 0; i < 6; i++) qac[i] = sp[(2 * NU + 13 + NU + i) * st];
        int steps = A.step_count[e], epi = A.episode[e];
        if (do_reset) {
            float x, y;
            if (A.cube_xy != nullptr) { x = A.cube_xy[2 * e]; y = A.cube_xy[2 * e + 1]; }
            else gg::spawn_draw(A.M, A.seed, gid, (uint32_t)epi, x, y);
            cp = gg::mk(x, y, A.M.spawn_z); cq = gg::ldq(A.M.spawn_quat);
            cv = gg::mk(0.f, 0.f, 0.f); cw = gg::mk(0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 6; i++) qac[i] = 0.f;
            epi += 1; steps = 0;
        }
        bool commit = live;
        if (run_sub) {
            X.cube[0][lane] = cp.x; X.cube[1][lane] = cp.y; X.cube[2][lane] = cp.z;
            X.cube[3][lane] = cq.w; X.cube[4][lane] = cq.x; X.cube[5][lane] = cq.y; X.cube[6][lane] = cq.z;
            X.cube[7][lane] = cv.x; X.cube[8][lane] = cv.y; X.cube[9][lane] = cv.z;
            X.cube[10][lane] = cw.x; X.cube[11][lane] = cw.y; X.cube[12][lane] = cw.z;
#pragma unroll
            for (int i = 0; i < 6; i++) X.cube[13 + i][lane] = qac[i];
            bar_arrive(8, 96 * SUB);
            float Mc[6], as_c[6], qd_c[6], a_c[6];
            gg::cube_free_acc(A.M, cv, cw, Mc, as_c, qd_c);
            gg::StepInfo ic; ic.ncon = 0; ic.iters = 0; ic.grad = 0.f;
#pragma unroll
            for (int i = 0; i < 6; i++) a_c[i] = as_c[i];
            if (commit) {
                // speculative: the broad phase (end of the RNEA chain) has not reported yet; coupled envs drop this result
                gg::CubeConStrided cs;                        // per-contact temporaries behind MAIN's slice of the team buffers
                cs.b = reinterpret_cast<float*>(TB) + gg::LinkStoreStrided<NA>::FLOATS * 32 + lane;
                gg::cube_block_solve(A.M, cp, gg::qmat(cq), Mc, as_c, qd_c, qac, a_c, A.iterations, A.tol, cs, ic);
            }
            bar_sync(1, 96 * SUB);
            const bool coupled = X.coupled[lane] != 0;
            {
                const unsigned cm = __ballot_sync(0xffffffffu, coupled);
                const int ncpl = __popc(cm);
                if (ncpl > 0) {                               // solver team 2 (see MAIN for the protocol)
                    bar_sync(pbar, 96);
                    team_queue<NA, NF, FH>(A, TW, TB, X, cm, ncpl, 2, lane, blk);
                    bar_sync(pbar, 96);
                }
            }
            bar_sync(2, 64 * SUB);
            if (coupled) {
#pragma unroll
                for (int i = 0; i < 6; i++) a_c[i] = X.aout[NU + i][lane];
                ic.ncon = X.ncon[lane]; ic.iters = 0;
            }
            const int it = ic.iters > X.iters[lane] ? ic.iters : X.iters[lane];
            if (active) {
                gg::integrate_cube(A.M, cp, cq, cv, cw, qac, a_c);
                st_iter = (float)it; st_con = (float)ic.ncon;
            }
        }
        {
            float acc = cp.x + cp.y + cp.z + cq.w + cq.x + cq.y + cq.z + cv.x + cv.y + cv.z + cw.x + cw.y + cw.z;
            X.bad[1][lane] = (run_sub && active && commit && !isfinite(acc)) ? 1 : 0;
        }
        __threadfence_block();
        bar_sync(3, 64 * SUB);
        const bool bad = (X.bad[0][lane] | X.bad[1][lane]) != 0;
        if (__any_sync(0xffffffffu, bad)) {
            if (bad) {
                float x, y;
                gg::spawn_draw(A.M, A.seed, gid, (uint32_t)epi, x, y);
                cp = gg::mk(x, y, A.M.spawn_z); cq = gg::ldq(A.M.spawn_quat);
                cv = gg::mk(0.f, 0.f, 0.f); cw = gg::mk(0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 6; i++) qac[i] = 0.f;
                epi += 1; steps = 0; st_nan = 1.f;
                do_reset = true;
            }
            bar_sync(pbar, 64);
        }
        const gg::V3 ep = gg::mk(X.ep[0][lane], X.ep[1][lane], X.ep[2][lane]);
        if (A.mode == 0 && !do_reset) { steps += 1; st_steps = 1.f; }
        const gg::V3 d = ep - cp;
        const bool succ = cp.z > A.M.success_z;
        const float rew = succ ? 1.f : 0.f;
        const bool trunc = A.time_limit && steps >= A.M.horizon;
        const float dn = sqrtf(gg::dot(d, d));
        if (full && A.vec_out) {
            float* r = X.oe + lane * GG_ENV_STATE_DIM;
            r[0] = cp.x; r[1] = cp.y; r[2] = cp.z; r[3] = cq.w; r[4] = cq.x; r[5] = cq.y; r[6] = cq.z;
            r[7] = d.x; r[8] = d.y; r[9] = d.z; r[10] = dn;
            __syncwarp();
            store_slab<32 * GG_ENV_STATE_DIM / 4>(X.oe, A.obs_env + (size_t)blk * 32 * GG_ENV_STATE_DIM,
                                                  A.obs_env_host ? A.obs_env_host + (size_t)blk * 32 * GG_ENV_STATE_DIM : nullptr, lane);
        } else if (commit) {
            float* oe = A.obs_env + (size_t)e * GG_ENV_STATE_DIM;
            oe[0] = cp.x; oe[1] = cp.y; oe[2] = cp.z; oe[3] = cq.w; oe[4] = cq.x; oe[5] = cq.y; oe[6] = cq.z;
            oe[7] = d.x; oe[8] = d.y; oe[9] = d.z; oe[10] = dn;
            if (A.obs_env_host) {
                float* oh = A.obs_env_host + (size_t)e * GG_ENV_STATE_DIM;
                oh[0] = cp.x; oh[1] = cp.y; oh[2] = cp.z; oh[3] = cq.w; oh[4] = cq.x; oh[5] = cq.y; oh[6] = cq.z;
                oh[7] = d.x; oh[8] = d.y; oh[9] = d.z; oh[10] = dn;
            }
        }
        if (commit) {
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
                if (A.mode == 0 && (succ || trunc)) { st_done = 1.f; st_succ = succ ? 1.f : 0.f; st_len = (float)steps; }
                if (A.mode == 0) st_ret = rew;
            }
        }
        if (!commit) { st_steps = st_done = st_succ = st_ret = st_len = st_nan = st_iter = st_con = 0.f; }
    }
    if (A.stats != nullptr && A.mode == 0) {
        float v[8] = {st_steps, st_done, st_succ, st_ret, st_len, st_nan, st_iter, st_con};
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == k) mine = x;
        }
        if (lane < 8 && mine != 0.f) atomicAdd(&A.stats[lane], (double)mine);
    }
}

#include "gg_drain.cuh"

// ------------------------------------------------------------------------------------------------
// step_mono_kernel: ONE THREAD = the whole env.step of one env (no role split).  The role kernel above shortens the dependent chain
// of a launch that cannot fill the GPU (<= ~32 k envs: 2.6 warps per scheduler) at the price of 41 % more instructions (three roles
// walk the same kinematic chain); with >= 7 env-warps per scheduler the GPU is instruction-supply bound and the cheaper instruction
// count wins (profiles/r01_notes.md v5: 749 M vs 504 M env-steps/s at 131072 envs; r02_notes.md for this version).  Same device
// functions, same arithmetic per env, same coupled-env list / drain; launch_step picks the kernel by env count.
// ------------------------------------------------------------------------------------------------
#ifndef GG_MONO_THREADS
#define GG_MONO_THREADS 64
#endif
constexpr int kMonoThreads = GG_MONO_THREADS;
#ifndef GG_MONO_MINBLOCKS
#define GG_MONO_MINBLOCKS (256 / GG_MONO_THREADS)
#endif
// -DGG_MONO_SYNC (experiment): block-wide barriers between the stages of the thread-per-env kernel keep the warps of a block at the same
// point of the instruction stream
#ifdef GG_MONO_SYNC
#define MONO_SYNC() __syncthreads()
#else
#define MONO_SYNC()
#endif
// Shared memory of one warp of the thread-per-env kernel: until the warp starts draining the coupled list its team buffers double as the
// [element][lane] scratch of the rolled loops (per-link S / I of the CRBA, then per-contact rows of the cube block) -- in local memory these
// arrays were the kernel's top stall (ncu: 2.0 of 6.1 cycles per issue waiting on L1TEX, 113 MB of live stack against a 126 MB L2,
// 223 MB of DRAM write-back per launch at 131072 envs; profiles/r02_step_mono_details.txt)
template <int NA, int NF> struct MonoWarp {
    static constexpr int SCRATCH = (gg::DynStoreStrided<NA>::FLOATS > gg::CubeConStrided::FLOATS ? gg::DynStoreStrided<NA>::FLOATS : gg::CubeConStrided::FLOATS) * 32;
    struct Team { gg::TeamWork<NA, NF> TW; gg::TeamBuf<NA + NF + 6> TB; };
    union { Team t; float scratch[SCRATCH]; };
};
template <int NA, int NF> struct MonoSmem { MonoWarp<NA, NF> w[kMonoThreads / 32]; };
constexpr int kMonoEnvsPerSm = 32 * GG_MINBLOCKS;   // more groups of 32 envs than ONE wave of role blocks holds (4 per SM: 18944 envs on a B200): the role
                                                    // kernel would need a second wave (16384 envs 36 us, 20480 envs 54 us), the thread-per-env kernel
                                                    // does not (42.6 / 44.4 us) -- profiles/r02_notes.md

template <int NA, int NF, bool FH>
__global__ void __launch_bounds__(kMonoThreads, GG_MONO_MINBLOCKS) step_mono_kernel(const __grid_constant__ StepArgs A) {
    constexpr int NU = NA + NF, NV = NU + 6, AD = 7 + NF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    MonoWarp<NA, NF>& SW = reinterpret_cast<MonoSmem<NA, NF>*>(smem_raw)->w[w];
    const int e0 = blockIdx.x * kMonoThreads + threadIdx.x;
    const bool live = e0 < A.n;
    const int e = live ? e0 : A.n - 1;
    const uint32_t gid = (uint32_t)(A.first_env + e);
    const bool run_sub = A.mode == 0 || (A.mode == 1 && A.M.settle_steps > 0);
    bool do_reset = reset_decision(A, e, gid);
    const bool active = do_reset ? (A.M.settle_steps > 0) : (A.mode == 0);
    const float* sp = A.state + e;
    const size_t st = (size_t)A.stride;
    float q[NU], qd[NU], qa[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) {
        const float lq = sp[i * st], lqd = sp[(NU + i) * st], lqa = sp[(2 * NU + 13 + i) * st];
        q[i] = do_reset ? A.M.home_q[i] : lq; qd[i] = do_reset ? 0.f : lqd; qa[i] = do_reset ? 0.f : lqa;
    }
    const float* cpn = sp + (2 * NU) * st;
    gg::V3 cp = gg::mk(cpn[0], cpn[st], cpn[2 * st]);
    gg::Q4 cq = gg::mkq(cpn[3 * st], cpn[4 * st], cpn[5 * st], cpn[6 * st]);
    gg::V3 cv = gg::mk(cpn[7 * st], cpn[8 * st], cpn[9 * st]);
    gg::V3 cw = gg::mk(cpn[10 * st], cpn[11 * st], cpn[12 * st]);
    float qac[6];
#pragma unroll
    for (int i = 0; i < 6; i++) qac[i] = sp[(2 * NU + 13 + NU + i) * st];
    int steps = A.step_count[e], epi = A.episode[e];
    if (do_reset) {
        float x, y;
        if (A.cube_xy != nullptr) { x = A.cube_xy[2 * e]; y = A.cube_xy[2 * e + 1]; }
        else gg::spawn_draw(A.M, A.seed, gid, (uint32_t)epi, x, y);
        cp = gg::mk(x, y, A.M.spawn_z); cq = gg::ldq(A.M.spawn_quat);
        cv = gg::mk(0.f, 0.f, 0.f); cw = gg::mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 6; i++) qac[i] = 0.f;
        epi += 1; steps = 0;
    }
    cl::Tally T;
#pragma unroll
    for (int k = 0; k < 8; k++) T.v[k] = 0.f;
    bool coupled = false;
    if (run_sub) {
        gg::RobotDyn<NA, NF> D;
        {
            ActionFetch<NU> tf;
            tf.home = A.M.home_q; tf.slab = nullptr; tf.global = A.action + (size_t)e * NU;
            tf.staged = false; tf.use_action = A.mode == 0 && !do_reset;
            gg::DynStoreStrided<NA> rs;                     // S, I of every arm link stay in the scratch: the mass-matrix pass below re-uses them
            rs.b = SW.scratch + lane;
            gg::robot_rnea_t<NA, NF, FH>(A.M, A.C, q, qd, tf, D.fs, rs, &D);
            if (A.action_dev != nullptr && A.mode == 0 && live) {
#pragma unroll
                for (int i = 0; i < NU; i++) A.action_dev[(size_t)e * NU + i] = A.action[(size_t)e * NU + i];
            }
        }
        coupled = live && active && gg::pads_may_touch<NA, NF>(A.M, D, cp, cq);
        MONO_SYNC();
        {
            gg::DynStoreStrided<NA> rs;                     // no second chain walk: frames and inertias come from the bias-force pass
            rs.b = SW.scratch + lane;
            gg::robot_crba_backward<NA, NF, FH>(A.M, D, rs);
        }
        MONO_SYNC();
        float L[NU * (NU + 1) / 2], dinv[NU];
#pragma unroll
        for (int i = 0; i < NU * (NU + 1) / 2; i++) L[i] = D.Mt[i];
        gg::chol_packed<NU>(L, dinv);
        float lsg[NU], laref[NU], lD[NU];
        const int nlim = gg::joint_limit_rows<NU>(A.M, q, qd, lsg, laref, lD);
        float as[NU], a[NU];
        gg::chol_solve_packed<NU>(L, dinv, D.fs, as);
        float Mc[6], as_c[6], qd_c[6], a_c[6];
        gg::cube_free_acc(A.M, cv, cw, Mc, as_c, qd_c);
        gg::StepInfo ir, ic;
        ir.ncon = ic.ncon = 0; ir.iters = ic.iters = 0; ir.grad = ic.grad = 0.f;
#pragma unroll
        for (int i = 0; i < NU; i++) a[i] = as[i];
#pragma unroll
        for (int i = 0; i < 6; i++) a_c[i] = as_c[i];
        MONO_SYNC();
        if (live && !coupled) gg::robot_block_solve<NU>(D.Mt, L, dinv, as, nlim, lsg, laref, lD, qa, a, A.iterations, A.tol, ir);
        MONO_SYNC();
        if (live && !coupled) {
            gg::CubeConStrided cs;                          // the CRBA is done with the scratch
            cs.b = SW.scratch + lane;
            gg::cube_block_solve(A.M, cp, gg::qmat(cq), Mc, as_c, qd_c, qac, a_c, A.iterations, A.tol, cs, ic);
        }
        MONO_SYNC();
        const unsigned cm = __ballot_sync(0xffffffffu, coupled);
        if (cm != 0u) {
            if (coupled) {
                constexpr int TWF = sizeof(gg::TeamWork<NA, NF>) / sizeof(float);
                gg::TeamWork<NA, NF>& W = *reinterpret_cast<gg::TeamWork<NA, NF>*>(A.team_scratch + (size_t)e0 * TWF);
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
                W.cube[0] = cp.x; W.cube[1] = cp.y; W.cube[2] = cp.z; W.cube[3] = cq.w; W.cube[4] = cq.x; W.cube[5] = cq.y; W.cube[6] = cq.z;
                W.cube[7] = cv.x; W.cube[8] = cv.y; W.cube[9] = cv.z; W.cube[10] = cw.x; W.cube[11] = cw.y; W.cube[12] = cw.z;
#pragma unroll
                for (int i = 0; i < 6; i++) W.cube[13 + i] = qac[i];
            }
            cl::list_append(A, cm, coupled, e0, lane);
        }
        if (active && !coupled) {
            gg::integrate_robot<NU>(A.M, q, qd, qa, a);
            gg::integrate_cube(A.M, cp, cq, cv, cw, qac, a_c);
            T.v[6] = (float)(ic.iters > ir.iters ? ic.iters : ir.iters); T.v[7] = (float)ic.ncon;
        }
    }
    const bool commit = live && !coupled;
    {
        float acc = cp.x + cp.y + cp.z + cq.w + cq.x + cq.y + cq.z + cv.x + cv.y + cv.z + cw.x + cw.y + cw.z;
#pragma unroll
        for (int i = 0; i < NU; i++) acc += q[i] + qd[i];
        if (run_sub && active && commit && !isfinite(acc)) {          // diverged env: re-spawn
            float x, y;
            gg::spawn_draw(A.M, A.seed, gid, (uint32_t)epi, x, y);
            cp = gg::mk(x, y, A.M.spawn_z); cq = gg::ldq(A.M.spawn_quat);
            cv = gg::mk(0.f, 0.f, 0.f); cw = gg::mk(0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 6; i++) qac[i] = 0.f;
            // NOTE: fresh array selected element-wise (in-place conditional overwrite of q[] was miscompiled by nvcc 12.9, r01 notes)
            float qh[NU];
#pragma unroll
            for (int i = 0; i < NU; i++) qh[i] = A.M.home_q[i];
#pragma unroll
            for (int i = 0; i < NU; i++) { q[i] = qh[i]; qd[i] = 0.f; qa[i] = 0.f; }
            epi += 1; steps = 0; T.v[5] = 1.f;
            do_reset = true;
        }
    }
    gg::V3 ep; gg::Q4 eq;
    gg::robot_eef<NA, NF>(A.M, q, ep, eq);
    if (commit) {
        const gg::V3 d = ep - cp;
        float oa[AD], oe[GG_ENV_STATE_DIM];
        oa[0] = ep.x; oa[1] = ep.y; oa[2] = ep.z; oa[3] = eq.w; oa[4] = eq.x; oa[5] = eq.y; oa[6] = eq.z;
#pragma unroll
        for (int f = 0; f < NF; f++) oa[7 + f] = q[NA + f];
        oe[0] = cp.x; oe[1] = cp.y; oe[2] = cp.z; oe[3] = cq.w; oe[4] = cq.x; oe[5] = cq.y; oe[6] = cq.z;
        oe[7] = d.x; oe[8] = d.y; oe[9] = d.z; oe[10] = sqrtf(gg::dot(d, d));
#pragma unroll
        for (int i = 0; i < AD; i++) { A.obs_agent[(size_t)e * AD + i] = oa[i]; if (A.obs_agent_host) A.obs_agent_host[(size_t)e * AD + i] = oa[i]; }
#pragma unroll
        for (int i = 0; i < GG_ENV_STATE_DIM; i++) {
            A.obs_env[(size_t)e * GG_ENV_STATE_DIM + i] = oe[i];
            if (A.obs_env_host) A.obs_env_host[(size_t)e * GG_ENV_STATE_DIM + i] = oe[i];
        }
        if (A.mode != 2 && (active || do_reset)) {
            float* wp = A.state + e;
#pragma unroll
            for (int i = 0; i < NU; i++) { wp[i * st] = q[i]; wp[(NU + i) * st] = qd[i]; wp[(2 * NU + 13 + i) * st] = qa[i]; }
        }
        cl::commit_cube<NU>(A, e, cp, cq, cv, cw, qac, steps, epi, do_reset, active, T);
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
    // ---- coupled envs of the whole launch (gg_drain.cuh)
    __syncwarp();                                           // every lane is done with the scratch view
    if (run_sub && !A.defer_drain) cl::drain_all<NA, NF, FH>(A, lane, SW.t.TB, SW.t.TW);
}

// ------------------------------------------------------------------------------------------------
// drain_kernel: the coupled envs the thread-per-env kernel listed, one WARP per env, as a kernel of its own (launched right behind
// step_mono_kernel on the same stream).  The warp-per-env solve is latency bound per warp, so its throughput is the number of resident
// warps: inside step_mono_kernel that is 8 per SM (254 registers of the thread-per-env code), here 16 (<= 128 registers, 12 KB of
// shared memory per warp) -- scripted grasp at 131072 envs: 22.9 M env-steps/s drained in-kernel, 30.4 M with this kernel, 32.7 M with the role
// kernel's block-local teams.  Opt-in (GG_SPLIT_DRAIN=1): an empty list costs ~5 us per step, and launch_step switches to the role kernel
// anyway once more than a few per cent of the envs are coupled.
// ------------------------------------------------------------------------------------------------
constexpr int kDrainThreads = 128;
#ifndef GG_DRAIN_MINBLOCKS
#define GG_DRAIN_MINBLOCKS 4
#endif
template <int NA, int NF, bool FH>
__global__ void __launch_bounds__(kDrainThreads, GG_DRAIN_MINBLOCKS) drain_kernel(const __grid_constant__ StepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename MonoWarp<NA, NF>::Team Team;
    Team& T = reinterpret_cast<Team*>(smem_raw)[threadIdx.x >> 5];
    cl::drain_all<NA, NF, FH>(A, threadIdx.x & 31, T.TB, T.TW);
}

#ifdef GG_WITH_RS_KERNEL
#include "gg_step_rs.cuh"
#endif

// env-major (n, nq) / (n, nv) copies <-> SoA state
template <int NU>
__global__ void state_copy_kernel(float* state, int stride, int n, float* q, float* qd, float* qacc, int to_soa) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    constexpr int NQ = NU + 7, NV = NU + 6;
    gg::EnvState<NU> s;
    if (to_soa) {
        load_state<NU>(state, stride, e, s);
        if (q) {
            const float* p = q + (size_t)e * NQ;
#pragma unroll
            for (int i = 0; i < NU; i++) s.q[i] = p[i];
            s.cp = gg::mk(p[NU], p[NU + 1], p[NU + 2]);
            s.cq = gg::mkq(p[NU + 3], p[NU + 4], p[NU + 5], p[NU + 6]);
        }
        if (qd) {
            const float* p = qd + (size_t)e * NV;
#pragma unroll
            for (int i = 0; i < NU; i++) s.qd[i] = p[i];
            s.cv = gg::mk(p[NU], p[NU + 1], p[NU + 2]);
            s.cw = gg::mk(p[NU + 3], p[NU + 4], p[NU + 5]);
        }
        if (qacc) {
            const float* p = qacc + (size_t)e * NV;
#pragma unroll
            for (int i = 0; i < NV; i++) s.qacc[i] = p[i];
        }
        store_state<NU>(state, stride, e, s);
    } else {
        load_state<NU>(state, stride, e, s);
        if (q) {
            float* p = q + (size_t)e * NQ;
#pragma unroll
            for (int i = 0; i < NU; i++) p[i] = s.q[i];
            p[NU] = s.cp.x; p[NU + 1] = s.cp.y; p[NU + 2] = s.cp.z;
            p[NU + 3] = s.cq.w; p[NU + 4] = s.cq.x; p[NU + 5] = s.cq.y; p[NU + 6] = s.cq.z;
        }
        if (qd) {
            float* p = qd + (size_t)e * NV;
#pragma unroll
            for (int i = 0; i < NU; i++) p[i] = s.qd[i];
            p[NU] = s.cv.x; p[NU + 1] = s.cv.y; p[NU + 2] = s.cv.z;
            p[NU + 3] = s.cw.x; p[NU + 4] = s.cw.y; p[NU + 5] = s.cw.z;
        }
        if (qacc) {
            float* p = qacc + (size_t)e * NV;
#pragma unroll
            for (int i = 0; i < NV; i++) p[i] = s.qacc[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ik_kernel: one thread = one IK problem (row).  The chain walk and the 6x6 solve live in registers / a 168-byte local array;
// rows of a warp run in lock step until every row of the warp has converged (gg_ik.cuh).  Per row: read 36 B (q) + 28 B (target),
// write 36 B + 12 B -- latency bound like the step kernel.
// ------------------------------------------------------------------------------------------------
struct IkArgs {
    GgModel M; GgIkConfig K;
    int n_rows, stride;
    const float* state; const int32_t* envs_idx;
    const float* tpos; const float* tquat; const float* init_q;
    float* q_out; float* err_out; int32_t* iters_out;
};

template <int NA, int NF>
__global__ void __launch_bounds__(64) ik_kernel(const __grid_constant__ IkArgs A) {
    constexpr int NU = NA + NF;
    const int r0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = r0 < A.n_rows;
    const int r = live ? r0 : A.n_rows - 1;              // dead lanes shadow the last row (the warp stays converged)
    const int e = A.envs_idx ? A.envs_idx[r] : r;
    float q[NU];
    if (A.init_q) {
#pragma unroll
        for (int i = 0; i < NU; i++) q[i] = A.init_q[(size_t)r * NU + i];
    } else {
#pragma unroll
        for (int i = 0; i < NU; i++) q[i] = A.state[(size_t)i * A.stride + e];
    }
    gg::V3 tp = gg::mk(0.f, 0.f, 0.f);
    gg::Q4 tq = gg::mkq(1.f, 0.f, 0.f, 0.f);
    if (A.tpos) tp = gg::ld3(A.tpos + (size_t)r * 3);
    if (A.tquat) tq = gg::ldq(A.tquat + (size_t)r * 4);
    float perr, rerr;
    const int it = gg::env_ik<NA>(A.M, A.K, q, tp, tq, perr, rerr);
    if (!live) return;
#pragma unroll
    for (int i = 0; i < NU; i++) A.q_out[(size_t)r * NU + i] = q[i];
    if (A.err_out) { A.err_out[2 * r] = perr; A.err_out[2 * r + 1] = rerr; }
    if (A.iters_out) A.iters_out[r] = it;
}

// Re-sampling pass (GgIkConfig::max_samples > 1): ONE WARP per row.  A warp whose row was solved by ik_kernel leaves at once; otherwise
// its 32 lanes run attempts s = 1 + lane, 33 + lane, ... in lock step (independent random starts, ik_sample_init) and the warp keeps
// the successful attempt with the LOWEST sample index -- the one a sequential loop over the samples would have returned -- or, when
// no attempt meets the tolerances, the best-ranked one (attempt 0 included).  Rows that cannot be reached (3 % of the hover targets of the
// reference expert) cost 2 rounds x max_iters updates in ONE warp instead of holding 31 solved rows of their ik_kernel warp hostage.
template <int NA, int NF>
__global__ void __launch_bounds__(128) ik_resample_kernel(const __grid_constant__ IkArgs A) {
    constexpr int NU = NA + NF;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= A.n_rows) return;
    float p0 = A.err_out[2 * r], r0 = A.err_out[2 * r + 1];
    if (p0 < A.K.pos_tol && r0 < A.K.rot_tol) return;                       // warp-uniform
    gg::V3 tp = gg::mk(0.f, 0.f, 0.f);
    gg::Q4 tq = gg::mkq(1.f, 0.f, 0.f, 0.f);
    if (A.tpos) tp = gg::ld3(A.tpos + (size_t)r * 3);
    if (A.tquat) tq = gg::ldq(A.tquat + (size_t)r * 4);
    float qbest[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) qbest[i] = A.q_out[(size_t)r * NU + i];    // attempt 0 (fingers are never touched)
    float best = gg::ik_score(A.K, p0, r0), bp = p0, br = r0;
    int extra = 0;
    bool solved = false;
    for (int base = 1; base < A.K.max_samples && !solved; base += 32) {
        const int s = base + lane;
        const bool mine = s < A.K.max_samples;
        float q[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) q[i] = qbest[i];
        if (mine) gg::ik_sample_init<NA>(A.M, A.K.seed_lo, A.K.seed_hi, (uint32_t)r, (uint32_t)s, q);
        float pe, re;
        const int it = gg::env_ik<NA>(A.M, A.K, q, tp, tq, pe, re);         // lanes beyond max_samples re-run the best start (harmless)
        extra += it;
        const bool ok = mine && pe < A.K.pos_tol && re < A.K.rot_tol;
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        int src;
        if (okm != 0u) { src = __ffs(okm) - 1; solved = true; }
        else {
            // no success in this round: lane with the smallest score (ties: lowest sample index), if it beats the best so far
            float sc = mine ? gg::ik_score(A.K, pe, re) : 3.0e38f;
            float m = sc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const unsigned mm = __ballot_sync(0xffffffffu, sc == m);
            src = m < best ? __ffs(mm) - 1 : -1;
        }
        if (src >= 0) {
#pragma unroll
            for (int i = 0; i < NA; i++) qbest[i] = __shfl_sync(0xffffffffu, q[i], src);
            bp = __shfl_sync(0xffffffffu, pe, src); br = __shfl_sync(0xffffffffu, re, src);
            best = gg::ik_score(A.K, bp, br);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NA; i++) A.q_out[(size_t)r * NU + i] = qbest[i];
        A.err_out[2 * r] = bp; A.err_out[2 * r + 1] = br;
        if (A.iters_out) A.iters_out[r] += extra;
    }
}

// ------------------------------------------------------------------------------------------------
// Trajectory store.  Pure data movement, HBM bound:
//   record_kernel : frame t <- [agent_pos | environment_state | action | reward]; one thread per float, 4 B read + 4 B written,
//                   consecutive threads write consecutive addresses (the (n, F) frame slab is contiguous)
//   select_kernel : one block; ordered compaction of the envs whose reward was ever > 0 (block-wide exclusive scan per 1024 envs)
//   export_kernel : out[slot][t][:] <- traj[t][env][:]; block (x, env) copies a run of (t, c) pairs: reads are F-float runs,
//                   writes are fully contiguous
// ------------------------------------------------------------------------------------------------
__global__ void record_kernel(float* __restrict__ frame, int32_t* __restrict__ ever, int n, int AD, int ED, int NU, const float* __restrict__ agent,
                              const float* __restrict__ envs, const float* __restrict__ action, const float* __restrict__ reward) {
    const int F = AD + ED + NU + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * F) return;
    const int e = i / F, c = i - e * F;
    float v;
    if (c < AD) v = agent[(size_t)e * AD + c];
    else if (c < AD + ED) v = envs[(size_t)e * ED + (c - AD)];
    else if (c < F - 1) v = action[(size_t)e * NU + (c - AD - ED)];
    else { v = reward[e]; if (v > 0.f) ever[e] = 1; }
    frame[i] = v;
}

__global__ void __launch_bounds__(1024) select_kernel(const int32_t* __restrict__ ever, int n, int32_t* __restrict__ slot,
                                                      int32_t* __restrict__ env_ids, int32_t* __restrict__ count) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int e0 = 0; e0 < n; e0 += 1024) {
        const int e = e0 + threadIdx.x;
        const int f = (e < n && ever[e] != 0) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wsum[w] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int k = 0; k < w; k++) off += wsum[k];
        const int pos = off + __popc(m & ((1u << lane) - 1u));
        if (e < n) slot[e] = f ? pos : -1;
        if (f) env_ids[pos] = e;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 32; k++) t += wsum[k]; base_s += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base_s;
}

__global__ void export_kernel(const float* __restrict__ traj, const int32_t* __restrict__ slot, int T, int n, int F, float* __restrict__ out) {
    const int e = blockIdx.y;
    const int s = slot[e];
    if (s < 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * F) return;
    const int t = i / F, c = i - t * F;
    out[(size_t)s * T * F + i] = traj[((size_t)t * n + e) * F + c];
}

__global__ void stats_kernel(double* stats, double* out, int clear) {
    int i = threadIdx.x;
    if (i < 8) { out[i] = stats[i]; if (clear) stats[i] = 0.0; }
}

}  // namespace

struct GgSim {
    GgModel model;
    GgConfig cfg;
    int n, device, stride, kind;        // kind: 0 = <7,2,slide>, 1 = <5,1,hinge>
    float* state; int32_t* step_count; int32_t* episode;
    uint8_t* done; double* stats;       // library-owned (allocated once in gg_create)
    float* team_scratch;                // library-owned: published dynamics of coupled envs (step_kernel solver teams)
    float* mbox; uint32_t* mflag; int* ctl; int* clist; uint32_t* cready;   // library-owned: role-per-SM kernel mailboxes / coupled list
    uint32_t epoch; int n_sm; int use_blocks; int use_mono; int sub; int roomy; int split_drain; int kernel_forced; int64_t list_budget;   // launch epoch of the mailbox flags; SM count; 1 = round-1 block kernel (GG_STEP_KERNEL=blocks)
    int32_t* slot;                      // library-owned, allocated on first gg_export_successful: env -> compacted episode slot
    uint32_t step_index;
    GgPinCache pins;                    // address ranges already verified as pinned host memory (gg_step_host)
    // adaptive kernel choice above one wave of role blocks: lagged fraction of coupled envs (device counter copied to pinned memory every
    // 16 steps, read one window later -- never waited for)
    int* coupled_host; cudaEvent_t coupled_ev; int coupled_pending; int coupled_prev; uint32_t adapt_tick; float coupled_frac; int prefer_roles;
};

// Is `p` page-locked host memory the device can address (cudaHostAlloc / cudaHostRegister)?  The verdict is cached per ALLOCATION
// RANGE (driver attribute RANGE_START_ADDR / RANGE_SIZE, resolved lazily so that the library loads without a driver): callers that
// hand out a fresh slot of one pinned batch every step (gym_genesis_b200.sim fresh outputs) pay no driver query in the steady
// state.  A cached range is re-verified every 1024 hits, so a range that was freed and re-used by pageable memory is not trusted
// for long (the per-pointer cache of round 1 was never invalidated).
bool gg_internal_pinned(GgPinCache* pc, const void* p) {
    const uintptr_t a = (uintptr_t)p;
    for (int i = 0; i < GgPinCache::kN; i++) {
        if (pc->size[i] != 0 && a >= pc->base[i] && a < pc->base[i] + pc->size[i]) {
            if (++pc->hits[i] < 1024) return true;
            pc->size[i] = 0;                              // age out: fall through to a fresh query
            break;
        }
    }
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost && at.type != cudaMemoryTypeManaged) return false;
    typedef int (*PtrAttrFn)(void*, int, unsigned long long);
    static PtrAttrFn fn = []() -> PtrAttrFn {
        void* h = dlopen("libcuda.so.1", RTLD_LAZY | RTLD_GLOBAL);
        return h ? (PtrAttrFn)dlsym(h, "cuPointerGetAttribute") : nullptr;
    }();
    unsigned long long start = 0;
    size_t size = 0;
    const unsigned long long dp = (unsigned long long)(at.devicePointer ? at.devicePointer : p);
    if (fn && fn(&start, 11 /* CU_POINTER_ATTRIBUTE_RANGE_START_ADDR */, dp) == 0 && fn(&size, 12 /* ..._RANGE_SIZE */, dp) == 0 && size != 0 &&
        a >= (uintptr_t)start && a < (uintptr_t)start + size) {
        pc->base[pc->next] = (uintptr_t)start; pc->size[pc->next] = size;
    } else {
        pc->base[pc->next] = a; pc->size[pc->next] = 1;  // range unknown: cache the exact pointer only
    }
    pc->hits[pc->next] = 0;
    pc->next = (pc->next + 1) % GgPinCache::kN;
    return true;
}

// hooks for the other translation units of the library (gg_stack.cu): shared error string and launch counter
int gg_internal_fail(const char* fmt, const char* detail) { return fail(fmt, detail); }
void gg_internal_count_launches(int k) { g_launches.fetch_add(k); }

// trajectory store shared by the pick and the stacking scenes (frame = [agent_pos | environment_state | action | reward])
int gg_internal_record(int device, int n, int AD, int ED, int NU, float* traj, int32_t* ever, int t, int horizon, const float* action,
                       const float* obs_agent, const float* obs_env, const float* reward, gg_stream_t stream) {
    if (!traj || !ever || !action || !obs_agent || !obs_env || !reward) return fail("gg_record: null argument%s");
    if (t < 0 || t >= horizon) return fail("gg_record: frame index outside the horizon%s");
    const int F = AD + ED + NU + 1;
    GgDeviceGuard dev_guard(device); GG_CUDA(dev_guard.err);
    const int total = n * F;
    record_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(traj + (size_t)t * n * F, ever, n, AD, ED, NU, obs_agent, obs_env, action, reward);
    g_launches.fetch_add(1);
    GG_CUDA(cudaGetLastError());
    return 0;
}
int gg_internal_export(int device, int n, int stride, int F, int32_t** slot, const float* traj, const int32_t* ever, int horizon, float* out,
                       int32_t* out_env_ids, int32_t* out_count, gg_stream_t stream) {
    if (!traj || !ever || !out || !out_env_ids || !out_count) return fail("gg_export_successful: null argument%s");
    if (horizon <= 0) return fail("gg_export_successful: horizon must be positive%s");
    GgDeviceGuard dev_guard(device); GG_CUDA(dev_guard.err);
    if (!*slot) GG_CUDA(cudaMalloc(slot, (size_t)stride * sizeof(int32_t)));   // first use only
    select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ever, n, *slot, out_env_ids, out_count);
    dim3 grid((horizon * F + 255) / 256, n);
    export_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(traj, *slot, horizon, n, F, out);
    g_launches.fetch_add(2);
    GG_CUDA(cudaGetLastError());
    return 0;
}

template <int SUB, int MB> static void launch_roles(int kind, int ngroups, const StepArgs& A, cudaStream_t stream) {
    const int grid = (ngroups + SUB - 1) / SUB;
    if (kind == 0) step_kernel<7, 2, false, SUB, MB><<<grid, kBlockThreads * SUB, SUB * sizeof(StepSmem<7, 2>), stream>>>(A);
    else step_kernel<5, 1, true, SUB, MB><<<grid, kBlockThreads * SUB, SUB * sizeof(StepSmem<5, 1>), stream>>>(A);
}
template <int SUB, int MB> static void prepare_roles() {
    cudaFuncSetAttribute((const void*)step_kernel<7, 2, false, SUB, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SUB * sizeof(StepSmem<7, 2>)));
    cudaFuncSetAttribute((const void*)step_kernel<5, 1, true, SUB, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SUB * sizeof(StepSmem<5, 1>)));
}

static int launch_step(GgSim* sim, StepArgs& A, gg_stream_t stream) {
    if (sim->state == nullptr) return fail("gg_bind_state has not been called%s");
    A.M = sim->model; gg::make_chain(A.M, A.C); A.n = sim->n; A.stride = sim->stride;
    A.state = sim->state; A.step_count = sim->step_count; A.episode = sim->episode; A.done = sim->done;
    A.iterations = sim->cfg.solver_iterations; A.tol = sim->cfg.solver_tolerance;
    A.auto_reset = sim->cfg.auto_reset; A.time_limit = sim->cfg.time_limit; A.reset_prob = sim->cfg.reset_prob;
    A.seed = sim->cfg.seed; A.first_env = sim->cfg.first_env_id; A.step_index = sim->step_index;
    A.stats = sim->stats; A.team_scratch = sim->team_scratch;
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
    A.vec_out = al16(A.obs_agent) && al16(A.obs_env) && al16(A.obs_agent_host) && al16(A.obs_env_host) && al16(A.action_dev);
    GgDeviceGuard dev_guard(sim->device); GG_CUDA(dev_guard.err);
    A.ctl = sim->ctl; A.clist = sim->clist; A.cready = sim->cready; A.list_cap = sim->stride;
    A.epoch = ++sim->epoch;
    if (A.epoch == 0) A.epoch = ++sim->epoch;
    // the list counters grow monotonically (gg_drain.cuh); every launch appends at most n_envs entries: zero them, in stream order,
    // long before they could reach 2^31
    sim->list_budget += (int64_t)sim->n;
    if (sim->list_budget > ((int64_t)1 << 30)) {
        GG_CUDA(cudaMemsetAsync(sim->ctl, 0, 16 * sizeof(int), (cudaStream_t)stream));
        sim->list_budget = 0;
    }
#ifdef GG_WITH_RS_KERNEL
    if (!sim->use_blocks) {
        // experimental (profiles/r02_notes.md): one persistent CTA per SM, one role per CTA
        const int G = sim->n_sm, ngroups = (sim->n + 31) / 32;
        A.mbox = sim->mbox; A.mflag = sim->mflag;
        if (ngroups <= rs::kWarps * (G / 3)) { A.n_rnea = G / 3; A.n_cube = G / 3; }
        else { A.n_rnea = (int)(G * 0.25f); A.n_cube = (int)(G * 0.36f); }
        A.n_main = G - A.n_rnea - A.n_cube;
        void* args[] = {(void*)&A};
        if (sim->kind == 0)
            GG_CUDA(cudaLaunchCooperativeKernel((const void*)rs::step_rs_kernel<7, 2, false>, dim3(G), dim3(rs::kThreads), args,
                                                rs::kWarps * rs::WarpScratch<7, 2>::bytes, (cudaStream_t)stream));
        else
            GG_CUDA(cudaLaunchCooperativeKernel((const void*)rs::step_rs_kernel<5, 1, true>, dim3(G), dim3(rs::kThreads), args,
                                                rs::kWarps * rs::WarpScratch<5, 1>::bytes, (cudaStream_t)stream));
        g_launches.fetch_add(1);
        GG_CUDA(cudaGetLastError());
        return 0;
    }
#endif
    bool mono = sim->use_mono != 0;
    if (mono && sim->kernel_forced < 0 && A.mode == 0) {
        // Above one wave of role blocks the thread-per-env kernel wins on free-space steps (131072 envs: 147 vs 259 us), but its coupled envs
        // are solved by 8 warps per SM instead of 12 (scripted grasp: 23 M vs 33 M env-steps/s).  Break-even is ~2 - 4 % of the envs coupled;
        // the choice follows the measured fraction with hysteresis.  Both kernels produce bit-identical states, so switching is seamless.
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if ((sim->adapt_tick++ & 15u) == 0u && cudaStreamIsCapturing((cudaStream_t)stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
            if (sim->coupled_pending && cudaEventQuery(sim->coupled_ev) == cudaSuccess) {
                const int cur = sim->coupled_host[0], delta = cur - sim->coupled_prev;
                if (delta >= 0) sim->coupled_frac = (float)delta / (16.f * (float)sim->n);     // < 0: the counters were just re-zeroed
                sim->coupled_prev = cur;
                sim->coupled_pending = 0;
                if (sim->coupled_frac > 0.04f) sim->prefer_roles = 1;
                else if (sim->coupled_frac < 0.02f) sim->prefer_roles = 0;
            }
            if (!sim->coupled_pending) {
                GG_CUDA(cudaMemcpyAsync(sim->coupled_host, sim->ctl + kCtlCoupled, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
                GG_CUDA(cudaEventRecord(sim->coupled_ev, (cudaStream_t)stream));
                sim->coupled_pending = 1;
            }
        }
        if (sim->prefer_roles) mono = false;
    }
    if (mono) {
        // throughput regime (>= ~7 env-warps per scheduler): thread-per-env kernel, fewest instructions per env
        const int grid = (sim->n + kMonoThreads - 1) / kMonoThreads;
        const bool run_sub = A.mode == 0 || (A.mode == 1 && A.M.settle_steps > 0);
        A.defer_drain = (sim->split_drain && run_sub) ? 1 : 0;
        if (sim->kind == 0) step_mono_kernel<7, 2, false><<<grid, kMonoThreads, sizeof(MonoSmem<7, 2>), (cudaStream_t)stream>>>(A);
        else step_mono_kernel<5, 1, true><<<grid, kMonoThreads, sizeof(MonoSmem<5, 1>), (cudaStream_t)stream>>>(A);
        if (A.defer_drain) {
            // persistent warps, one per listed env at a time; an empty list costs one look at the counter per warp
            const int dgrid = sim->n_sm * GG_DRAIN_MINBLOCKS;
            constexpr int dw = kDrainThreads / 32;
            if (sim->kind == 0) drain_kernel<7, 2, false><<<dgrid, kDrainThreads, dw * sizeof(MonoWarp<7, 2>::Team), (cudaStream_t)stream>>>(A);
            else drain_kernel<5, 1, true><<<dgrid, kDrainThreads, dw * sizeof(MonoWarp<5, 1>::Team), (cudaStream_t)stream>>>(A);
            g_launches.fetch_add(1);
        }
    } else {
        // latency regime: 32 envs x 3 role warps per block
        const int ngroups = (sim->n + 31) / 32;
        switch (sim->sub) {
        case 4: launch_roles<4, GG_MINBLOCKS>(sim->kind, ngroups, A, (cudaStream_t)stream); break;
        case 2: launch_roles<2, GG_MINBLOCKS>(sim->kind, ngroups, A, (cudaStream_t)stream); break;
        default:
            if (sim->roomy) launch_roles<1, 2>(sim->kind, ngroups, A, (cudaStream_t)stream);
            else launch_roles<1, GG_MINBLOCKS>(sim->kind, ngroups, A, (cudaStream_t)stream);
            break;
        }
    }
    g_launches.fetch_add(1);
    GG_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

const char* gg_last_error(void) { return g_err; }
int gg_version(void) { return 100; }
int64_t gg_launch_count(void) { return g_launches.load(); }

int gg_create(const GgModel* model, int n_envs, int device, GgSim** out) {
    if (!model || !out) return fail("gg_create: null argument%s");
    if (n_envs <= 0) return fail("gg_create: n_envs must be positive%s");
    int kind = -1;
    if (model->n_arm == 7 && model->n_fing == 2 && !model->fing_hinge) kind = 0;
    else if (model->n_arm == 5 && model->n_fing == 1 && model->fing_hinge) kind = 1;
    if (kind < 0) return fail("gg_create: unsupported articulation (need 7 hinges + 2 slides or 5 hinges + 1 hinge)%s");
    if (model->n_pads < 0 || model->n_pads > GG_MAX_PADS) return fail("gg_create: too many pads%s");
    if (model->power != 2.0f) return fail("gg_create: solimp power must be 2%s");
    int ndev = 0;
    GG_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("gg_create: bad device index%s");
    GgDeviceGuard dev_guard(device); GG_CUDA(dev_guard.err);
    GgSim* s = new (std::nothrow) GgSim();
    if (!s) return fail("gg_create: out of host memory%s");
    memset(s, 0, sizeof(*s));
    s->model = *model; s->n = n_envs; s->device = device; s->kind = kind;
    s->stride = (n_envs + 31) / 32 * 32;
    s->cfg.solver_iterations = 10; s->cfg.solver_tolerance = 1e-5f;
    cudaError_t e1 = cudaMalloc(&s->done, (size_t)s->stride);
    cudaError_t e2 = cudaMalloc(&s->stats, 8 * sizeof(double));
    const size_t twf = kind == 0 ? sizeof(gg::TeamWork<7, 2>) : sizeof(gg::TeamWork<5, 1>);
    cudaError_t e3 = cudaMalloc(&s->team_scratch, (size_t)s->stride * twf);
    const size_t nready = (size_t)s->stride + 8192;       // list slots (+ slack for the experimental kernel's over-claimed tickets)
    cudaError_t e6 = cudaMalloc(&s->ctl, 8192 * sizeof(int));
    cudaError_t e7 = cudaMalloc(&s->clist, nready * sizeof(int));
    cudaError_t e8 = cudaMalloc(&s->cready, nready * sizeof(uint32_t));
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e6 != cudaSuccess || e7 != cudaSuccess || e8 != cudaSuccess) {
        cudaFree(s->done); cudaFree(s->stats); cudaFree(s->team_scratch); cudaFree(s->ctl); cudaFree(s->clist); cudaFree(s->cready); delete s;
        return fail("gg_create: cudaMalloc failed%s");
    }
    cudaMemset(s->done, 0, (size_t)s->stride);
    cudaMemset(s->stats, 0, 8 * sizeof(double));
    cudaMemset(s->ctl, 0, 8192 * sizeof(int));
    cudaMemset(s->cready, 0, nready * sizeof(uint32_t));
    s->use_blocks = 1;
    cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, device);
    {
        // kernel choice by grid size (GG_STEP_KERNEL=mono / roles overrides it for A/B measurements): the role kernel while the launch
        // cannot fill the GPU with env-warps, the thread-per-env kernel above kMonoEnvsPerSm envs per SM
        const char* k = getenv("GG_STEP_KERNEL");
        s->use_mono = n_envs > kMonoEnvsPerSm * s->n_sm;
        s->kernel_forced = -1;
        if (k && strcmp(k, "mono") == 0) { s->use_mono = 1; s->kernel_forced = 1; }
        if (k && strcmp(k, "roles") == 0) { s->use_mono = 0; s->kernel_forced = 0; }
        s->coupled_pending = 0; s->coupled_prev = 0; s->adapt_tick = 0; s->coupled_frac = 0.f; s->prefer_roles = 0;
        s->coupled_host = nullptr; s->coupled_ev = nullptr;
        if (cudaHostAlloc((void**)&s->coupled_host, 2 * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->coupled_ev, cudaEventDisableTiming) != cudaSuccess) {
            s->kernel_forced = s->use_mono;          // no adaptive choice without the mirror
        } else s->coupled_host[0] = 0;
        // groups of 32 envs per block of the role kernel.  Wider blocks (GG_STEP_SUB=2 / 4: the same-role warps of a block meet at
        // SUB-times wider named barriers and run their instruction stream in step) were measured at 4096 .. 24576 envs and change
        // nothing (16384 envs: 35.4 / 35.6 / 36.1 us for 1 / 2 / 4, profiles/r02_notes.md), so one group per block stays the default
        s->sub = 1;
        const char* sb = getenv("GG_STEP_SUB");
        if (sb && (atoi(sb) == 1 || atoi(sb) == 2 || atoi(sb) == 4)) s->sub = atoi(sb);
        prepare_roles<1, GG_MINBLOCKS>(); prepare_roles<2, GG_MINBLOCKS>(); prepare_roles<4, GG_MINBLOCKS>(); prepare_roles<1, 2>();
        // launches of at most two role blocks per SM (9472 envs on 148 SMs) take the 246-register build of the role kernel
        s->roomy = (n_envs + 31) / 32 <= 2 * s->n_sm;
        const char* rm = getenv("GG_STEP_ROOMY");
        if (rm) s->roomy = atoi(rm) != 0 && (n_envs + 31) / 32 <= 2 * s->n_sm;
        cudaFuncSetAttribute((const void*)step_mono_kernel<7, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MonoSmem<7, 2>));
        cudaFuncSetAttribute((const void*)step_mono_kernel<5, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MonoSmem<5, 1>));
        cudaFuncSetAttribute((const void*)drain_kernel<7, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kDrainThreads / 32 * sizeof(MonoWarp<7, 2>::Team)));
        cudaFuncSetAttribute((const void*)drain_kernel<5, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kDrainThreads / 32 * sizeof(MonoWarp<5, 1>::Team)));
        // coupled envs of the thread-per-env kernel: drained inside step_mono_kernel (GG_SPLIT_DRAIN=1: by a kernel of their own, 16 instead
        // of 8 warps per SM -- 30 M instead of 23 M env-steps/s on the scripted grasp, but +5 us per step when the list is empty)
        s->split_drain = 0;
        const char* sd = getenv("GG_SPLIT_DRAIN");
        if (sd) s->split_drain = atoi(sd) != 0;
    }
#ifdef GG_WITH_RS_KERNEL
    {
        const size_t ngroups = (size_t)s->stride / 32;
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        const char* k = getenv("GG_STEP_KERNEL");
        if (k && strcmp(k, "rs") == 0 && coop && s->n_sm >= 3 &&
            cudaMalloc(&s->mbox, ngroups * rs::MB_ROWS * 32 * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&s->mflag, ngroups * 4 * sizeof(uint32_t)) == cudaSuccess) {
            cudaMemset(s->mflag, 0, ngroups * 4 * sizeof(uint32_t));
            cudaMemset(s->mbox, 0, ngroups * rs::MB_ROWS * 32 * sizeof(float));
            cudaFuncSetAttribute((const void*)rs::step_rs_kernel<7, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(rs::kWarps * rs::WarpScratch<7, 2>::bytes));
            cudaFuncSetAttribute((const void*)rs::step_rs_kernel<5, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(rs::kWarps * rs::WarpScratch<5, 1>::bytes));
            s->use_blocks = 0;
        }
    }
#endif
    *out = s;
    return 0;
}

int gg_destroy(GgSim* sim) {
    if (!sim) return 0;
    GgDeviceGuard dev_guard(sim->device);
    cudaFree(sim->done); cudaFree(sim->stats); cudaFree(sim->team_scratch); cudaFree(sim->slot);
    cudaFree(sim->mbox); cudaFree(sim->mflag); cudaFree(sim->ctl); cudaFree(sim->clist); cudaFree(sim->cready);
    if (sim->coupled_ev) { cudaEventSynchronize(sim->coupled_ev); cudaEventDestroy(sim->coupled_ev); }
    if (sim->coupled_host) cudaFreeHost(sim->coupled_host);
    delete sim;
    return 0;
}

int gg_configure(GgSim* sim, const GgConfig* cfg) {
    if (!sim || !cfg) return fail("gg_configure: null argument%s");
    if (cfg->solver_iterations < 1 || cfg->solver_iterations > 1000) return fail("gg_configure: solver_iterations out of range%s");
    if (!(cfg->reset_prob >= 0.f && cfg->reset_prob <= 1.f)) return fail("gg_configure: reset_prob out of range%s");
    sim->cfg = *cfg;
    return 0;
}

int gg_state_rows(const GgSim* sim) { int nu = sim->model.n_arm + sim->model.n_fing; return 3 * nu + 19; }
int gg_state_stride(const GgSim* sim) { return sim->stride; }

int gg_bind_state(GgSim* sim, float* state, int32_t* step_count, int32_t* episode) {
    if (!sim || !state || !step_count || !episode) return fail("gg_bind_state: null argument%s");
    sim->state = state; sim->step_count = step_count; sim->episode = episode;
    return 0;
}

int gg_step(GgSim* sim, const float* action, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated,
            uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream) {
    if (!sim || !action || !obs_agent || !obs_env) return fail("gg_step: null argument%s");
    StepArgs A;
    memset(&A, 0, sizeof(A));
    A.action = action; A.obs_agent = obs_agent; A.obs_env = obs_env; A.reward = reward;
    A.terminated = terminated; A.truncated = truncated; A.reset_mask = reset_mask; A.cube_xy = nullptr; A.mode = 0;
    A.stage_action = ((uintptr_t)action & 15u) == 0;
    int rc = launch_step(sim, A, stream);
    sim->step_index++;
    return rc;
}

static bool pinned_host_ptr(GgSim* sim, const void* p) { return gg_internal_pinned(&sim->pins, p); }

int gg_step_host(GgSim* sim, const GgHostIO* io, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated,
                 uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream) {
    if (!sim || !io || !io->action || !obs_agent || !obs_env) return fail("gg_step_host: null argument%s");
    if (!pinned_host_ptr(sim, io->action)) return fail("gg_step_host: action must be page-locked (pinned) host memory%s");
    if (io->reward && !pinned_host_ptr(sim, io->reward)) return fail("gg_step_host: reward mirror must be pinned host memory%s");
    if (io->terminated && !pinned_host_ptr(sim, io->terminated)) return fail("gg_step_host: terminated mirror must be pinned host memory%s");
    if (io->truncated && !pinned_host_ptr(sim, io->truncated)) return fail("gg_step_host: truncated mirror must be pinned host memory%s");
    if (io->obs_agent && !pinned_host_ptr(sim, io->obs_agent)) return fail("gg_step_host: obs_agent mirror must be pinned host memory%s");
    if (io->obs_env && !pinned_host_ptr(sim, io->obs_env)) return fail("gg_step_host: obs_env mirror must be pinned host memory%s");
    StepArgs A;
    memset(&A, 0, sizeof(A));
    A.action = io->action; A.obs_agent = obs_agent; A.obs_env = obs_env; A.reward = reward;
    A.terminated = terminated; A.truncated = truncated; A.reset_mask = reset_mask; A.cube_xy = nullptr; A.mode = 0;
    A.reward_host = io->reward; A.terminated_host = io->terminated; A.truncated_host = io->truncated;
    A.obs_agent_host = io->obs_agent; A.obs_env_host = io->obs_env; A.action_dev = io->action_dev;
    A.stage_action = ((uintptr_t)io->action & 15u) == 0;
    int rc = launch_step(sim, A, stream);
    sim->step_index++;
    if (rc != 0) return rc;
    if (io->wait) GG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int gg_reset(GgSim* sim, const uint8_t* mask, const float* cube_xy, float* obs_agent, float* obs_env, gg_stream_t stream) {
    if (!sim || !obs_agent || !obs_env) return fail("gg_reset: null argument%s");
    StepArgs A;
    memset(&A, 0, sizeof(A));
    A.obs_agent = obs_agent; A.obs_env = obs_env; A.reset_mask = mask; A.cube_xy = cube_xy; A.mode = 1;
    return launch_step(sim, A, stream);
}

int gg_observe(GgSim* sim, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated, gg_stream_t stream) {
    if (!sim || !obs_agent || !obs_env) return fail("gg_observe: null argument%s");
    StepArgs A;
    memset(&A, 0, sizeof(A));
    A.obs_agent = obs_agent; A.obs_env = obs_env; A.reward = reward; A.terminated = terminated; A.mode = 2;
    return launch_step(sim, A, stream);
}

static int state_copy(GgSim* sim, float* q, float* qd, float* qacc, int to_soa, gg_stream_t stream) {
    if (!sim || sim->state == nullptr) return fail("state copy: state not bound%s");
    GgDeviceGuard dev_guard(sim->device); GG_CUDA(dev_guard.err);
    const int grid = (sim->n + 127) / 128;
    if (sim->kind == 0) state_copy_kernel<9><<<grid, 128, 0, (cudaStream_t)stream>>>(sim->state, sim->stride, sim->n, q, qd, qacc, to_soa);
    else state_copy_kernel<6><<<grid, 128, 0, (cudaStream_t)stream>>>(sim->state, sim->stride, sim->n, q, qd, qacc, to_soa);
    g_launches.fetch_add(1);
    GG_CUDA(cudaGetLastError());
    return 0;
}
int gg_get_state(GgSim* sim, float* q, float* qd, float* qacc, gg_stream_t stream) { return state_copy(sim, q, qd, qacc, 0, stream); }
int gg_set_state(GgSim* sim, const float* q, const float* qd, const float* qacc, gg_stream_t stream) {
    return state_copy(sim, (float*)q, (float*)qd, (float*)qacc, 1, stream);
}

void gg_ik_defaults(GgIkConfig* cfg) {
    if (!cfg) return;
    cfg->max_iters = 20; cfg->damping = 0.01f; cfg->pos_tol = 5e-4f; cfg->rot_tol = 5e-3f; cfg->max_step = 0.5f;
    cfg->respect_limits = 1;
    cfg->max_samples = 1; cfg->seed_lo = 0u; cfg->seed_hi = 0u;
    for (int i = 0; i < 3; i++) { cfg->pos_mask[i] = 1.f; cfg->rot_mask[i] = 1.f; }
}

int gg_ik(GgSim* sim, int n_rows, const int32_t* envs_idx, const float* target_pos, const float* target_quat,
          const float* init_q, float* q_out, float* err_out, int32_t* iters_out, const GgIkConfig* cfg, gg_stream_t stream) {
    if (!sim || !q_out) return fail("gg_ik: null argument%s");
    if (!target_pos && !target_quat) return fail("gg_ik: at least one of target_pos / target_quat is required%s");
    if (n_rows <= 0) return fail("gg_ik: n_rows must be positive%s");
    if (!envs_idx && !init_q && n_rows > sim->n) return fail("gg_ik: n_rows exceeds n_envs (pass envs_idx or init_q)%s");
    if (!init_q && sim->state == nullptr) return fail("gg_ik: state not bound and no init_q given%s");
    IkArgs A;
    memset(&A, 0, sizeof(A));
    A.M = sim->model;
    if (cfg) A.K = *cfg; else gg_ik_defaults(&A.K);
    if (A.K.max_iters < 0 || A.K.max_iters > 10000) return fail("gg_ik: max_iters out of range%s");
    if (!(A.K.damping > 0.f)) return fail("gg_ik: damping must be positive%s");
    if (A.K.max_samples < 1) A.K.max_samples = 1;
    if (A.K.max_samples > 1 && !err_out) return fail("gg_ik: max_samples > 1 needs err_out (the re-sampling pass reads the first pass' errors)%s");
    if (!target_pos) A.K.pos_mask[0] = A.K.pos_mask[1] = A.K.pos_mask[2] = 0.f;
    if (!target_quat) A.K.rot_mask[0] = A.K.rot_mask[1] = A.K.rot_mask[2] = 0.f;
    A.n_rows = n_rows; A.stride = sim->stride; A.state = sim->state; A.envs_idx = envs_idx;
    A.tpos = target_pos; A.tquat = target_quat; A.init_q = init_q; A.q_out = q_out; A.err_out = err_out; A.iters_out = iters_out;
    GgDeviceGuard dev_guard(sim->device); GG_CUDA(dev_guard.err);
    const int grid = (n_rows + 63) / 64;
    if (sim->kind == 0) ik_kernel<7, 2><<<grid, 64, 0, (cudaStream_t)stream>>>(A);
    else ik_kernel<5, 1><<<grid, 64, 0, (cudaStream_t)stream>>>(A);
    g_launches.fetch_add(1);
    if (A.K.max_samples > 1) {
        const int g2 = (n_rows + 3) / 4;                    // one warp per row, four rows per block
        if (sim->kind == 0) ik_resample_kernel<7, 2><<<g2, 128, 0, (cudaStream_t)stream>>>(A);
        else ik_resample_kernel<5, 1><<<g2, 128, 0, (cudaStream_t)stream>>>(A);
        g_launches.fetch_add(1);
    }
    GG_CUDA(cudaGetLastError());
    return 0;
}

// developer hook (not in the public header): copies the first n ints of the control block to the host (GG_RS_TRACE timestamps)
// introspection for tests / tools: out = {size-based choice (1 = thread-per-env kernel), adaptive override towards the role kernel, lagged coupled fraction}
int gg_debug_dispatch(GgSim* sim, float* out) {
    if (!sim || !out) return fail("gg_debug_dispatch: bad argument%s");
    out[0] = (float)sim->use_mono; out[1] = (float)sim->prefer_roles; out[2] = sim->coupled_frac;
    return 0;
}

int gg_debug_ctl(GgSim* sim, int* out, int n) {
    if (!sim || !out || n > 8192) return fail("gg_debug_ctl: bad argument%s");
    GG_CUDA(cudaMemcpy(out, sim->ctl, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int gg_frame_dim(const GgSim* sim) { return sim ? (7 + sim->model.n_fing) + GG_ENV_STATE_DIM + (sim->model.n_arm + sim->model.n_fing) + 1 : -1; }

int gg_record(GgSim* sim, float* traj, int32_t* ever, int t, int horizon, const float* action, const float* obs_agent,
              const float* obs_env, const float* reward, gg_stream_t stream) {
    if (!sim) return fail("gg_record: null argument%s");
    return gg_internal_record(sim->device, sim->n, 7 + sim->model.n_fing, GG_ENV_STATE_DIM, sim->model.n_arm + sim->model.n_fing, traj, ever, t,
                              horizon, action, obs_agent, obs_env, reward, stream);
}

int gg_export_successful(GgSim* sim, const float* traj, const int32_t* ever, int horizon, float* out, int32_t* out_env_ids,
                         int32_t* out_count, gg_stream_t stream) {
    if (!sim) return fail("gg_export_successful: null argument%s");
    return gg_internal_export(sim->device, sim->n, sim->stride, gg_frame_dim(sim), &sim->slot, traj, ever, horizon, out, out_env_ids, out_count, stream);
}

int gg_stats(GgSim* sim, double* out8, int clear, gg_stream_t stream) {
    if (!sim || !out8) return fail("gg_stats: null argument%s");
    GgDeviceGuard dev_guard(sim->device); GG_CUDA(dev_guard.err);
    stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sim->stats, out8, clear);
    g_launches.fetch_add(1);
    GG_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
