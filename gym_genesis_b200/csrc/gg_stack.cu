// CUDA kernels (sm_100a) and C ABI of the multi-object (CubeStack) step path: include/gg_b200.h, gg_stack_*.
// Replaces, for gym_genesis/tasks/franka/cube_stack.py:119-216, the Genesis calls of reset / step / compute_reward / get_obs.
//
// One env.step = four launches on the caller's stream (per-env pipeline in gg_stack.cuh):
//   stack_prep_kernel    thread per env, 32 envs and three role warps per block.  Masked reset, robot dynamics, body-level broad phase (pads vs plane /
//                        cubes, cube vs cube); a robot whose pads can touch nothing is solved HERE in registers (robot block with
//                        joint-limit rows).  The StackIn record of
//                        each env (robot dynamics + object states) is transposed through shared memory and written env-major, so
//                        that both this kernel's stores and the solver's loads are coalesced.
//   stack_cubes_kernel   thread per (cube, env): the cubes that can touch nothing but the plane are solved in registers (6 x 6
//                        block, <= 4 plane contacts).
//   stack_solve_kernel   persistent warps drain the compacted list of envs that have bodies in reach of each other; one WARP per
//                        env, 4 warps per block, ~23 KB of dynamic shared memory per warp (contact table, pair stiffness, packed
//                        Hessian of the active dofs).  Accelerations go to a [dof][env] slab.
//   stack_finish_kernel  thread per env.  Integration, observations, reward / flags, counters, statistics.
// The records and the slab (22 MB at 16384 envs) stay L2 resident between the three kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "gg_b200.h"
#include "gg_internal.h"
#include "gg_stack.cuh"

int gg_internal_fail(const char* fmt, const char* detail);
void gg_internal_count_launches(int k);
int gg_internal_record(int device, int n, int AD, int ED, int NU, float* traj, int32_t* ever, int t, int horizon, const float* action,
                       const float* obs_agent, const float* obs_env, const float* reward, gg_stream_t stream);
int gg_internal_export(int device, int n, int stride, int F, int32_t** slot, const float* traj, const int32_t* ever, int horizon, float* out,
                       int32_t* out_env_ids, int32_t* out_count, gg_stream_t stream);

namespace {

#define SK_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return gg_internal_fail("CUDA error: %s", cudaGetErrorString(e_)); \
    } while (0)

constexpr unsigned kFullMask = 0xffffffffu;

struct WarpDev {
    int lane;
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(kFullMask, v, src); }
    __device__ __forceinline__ int shfli(int v, int src) const { return __shfl_sync(kFullMask, v, src); }
    __device__ __forceinline__ float sum(float v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
        return v;
    }
    __device__ __forceinline__ float max(float v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
        return v;
    }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(kFullMask, p) != 0; }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(kFullMask, p); }
    __device__ __forceinline__ int scan_incl(int v) const {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFullMask, v, o);
            if (lane >= o) v += u;
        }
        return v;
    }
};

struct StackArgs {
    GgModel M; gg::GgChain C;
    GgStackTask T;
    int n, stride;
    float* state; int32_t* step_count; int32_t* episode; uint8_t* done;
    const float* action; float* obs_agent; float* obs_env; float* reward;
    uint8_t* terminated; uint8_t* truncated;
    const uint8_t* reset_mask; const float* obj_xy;
    float* reward_host; uint8_t* terminated_host; uint8_t* truncated_host;   // nullable pinned-host mirrors (gg_stack_step_host)
    float* obs_agent_host; float* obs_env_host;   // nullable pinned-host observation mirrors
    float* action_dev;                            // nullable device copy of the consumed action batch
    int mode;                 // 0 = env.step, 1 = reset (+ settle) of masked envs, 2 = observe only
    int iterations; float tol;
    int auto_reset, time_limit; float reset_prob;
    uint64_t seed; int64_t first_env; uint32_t step_index;
    double* stats;
    float* work;              // n x sizeof(StackIn) / 4 floats, env-major
    float* acc;               // NV x stride accelerations, [dof][env]
    int32_t* ncon; int32_t* iters;
    uint8_t* flags;           // per env: bit 0 reset applied by this call, bit 1 a substep runs, bit 2 some bodies need the
                              // cooperative solve (otherwise stack_prep's register solves are the whole answer)
    int32_t* amask;           // per env: bodies of the cooperative solve (bit 0 robot, bit 1 + k cube k), written by stack_prep_kernel
    uint8_t* rob_it;          // per env: Newton iterations of the robot block's register solve
    uint8_t* iso_nc;          // [cube][env]: plane contacts / Newton iterations of the cubes solved in registers (stack_cubes_kernel)
    uint8_t* iso_it;
    int32_t* coop_list;       // envs that need the cooperative solve, appended by stack_prep_kernel (order irrelevant: envs are independent)
    int32_t* coop_ctr;        // [0] heavy-list length, [1] pop cursor of the persistent solver warps, [2] light-list length (the
                              // light list starts at coop_list + stride); cleared by stack_finish_kernel
};

// state rows: q [0, NU) | qd [NU, 2NU) | objects 13 K | warm start NU + 6 K
template <int NU, int K> struct Rows {
    static constexpr int Q = 0, QD = NU, OBJ = 2 * NU, WARM = 2 * NU + 13 * K;
};

__device__ __forceinline__ bool stack_reset_decision(const StackArgs& A, int e, uint32_t gid) {
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

template <int K>
__device__ __forceinline__ void spawn_objects(const StackArgs& A, int e, uint32_t gid, uint32_t epi, bool use_xy, gg::ObjState<K>& O) {
    float xy[K][2];
    if (use_xy && A.obj_xy != nullptr) {
#pragma unroll
        for (int k = 0; k < K; k++) { xy[k][0] = A.obj_xy[((size_t)e * K + k) * 2]; xy[k][1] = A.obj_xy[((size_t)e * K + k) * 2 + 1]; }
    } else {
        gg::stack_spawn_all(A.T, A.seed, gid, epi, K, xy);
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float x = xy[k][0], y = xy[k][1];
        O.p[k] = gg::mk(x, y, A.M.spawn_z); O.q[k] = gg::ldq(A.M.spawn_quat);
        O.v[k] = gg::mk(0.f, 0.f, 0.f); O.w[k] = gg::mk(0.f, 0.f, 0.f);
    }
}

template <int NU, int K>
__device__ __forceinline__ void load_objects(const float* sp, size_t st, gg::ObjState<K>& O) {
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float* c = sp + (Rows<NU, K>::OBJ + 13 * k) * st;
        O.p[k] = gg::mk(c[0], c[st], c[2 * st]);
        O.q[k] = gg::mkq(c[3 * st], c[4 * st], c[5 * st], c[6 * st]);
        O.v[k] = gg::mk(c[7 * st], c[8 * st], c[9 * st]);
        O.w[k] = gg::mk(c[10 * st], c[11 * st], c[12 * st]);
    }
}
template <int NU, int K>
__device__ __forceinline__ void store_objects(float* sp, size_t st, const gg::ObjState<K>& O) {
#pragma unroll
    for (int k = 0; k < K; k++) {
        float* c = sp + (Rows<NU, K>::OBJ + 13 * k) * st;
        c[0] = O.p[k].x; c[st] = O.p[k].y; c[2 * st] = O.p[k].z;
        c[3 * st] = O.q[k].w; c[4 * st] = O.q[k].x; c[5 * st] = O.q[k].y; c[6 * st] = O.q[k].z;
        c[7 * st] = O.v[k].x; c[8 * st] = O.v[k].y; c[9 * st] = O.v[k].z;
        c[10 * st] = O.w[k].x; c[11 * st] = O.w[k].y; c[12 * st] = O.w[k].z;
    }
}

// ------------------------------------------------------------------------------------------------------------------ prep
// Three role warps per 32 envs (lane = env), as in the CubePick step kernel: warp 0 ROBOT (CRBA, Cholesky, limit rows, robot-block
// solve), warp 1 RNEA (bias forces + PD torque of the same chain, handed over through shared memory and named barrier 1), warp 2
// CUBES (pad reach from its own walk of the pad-body chain, cube-cube reach, register solves of the isolated cubes, reset
// spawn).  The pieces are independent until the amask words are merged after one __syncthreads; running them side by side cuts
// the dependent chain of this latency-bound kernel and triples its warps per SM.
__device__ __forceinline__ void sk_bar_sync(int id, int nthreads) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sk_bar_arrive(int id, int nthreads) {
    __threadfence_block();
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int NU> struct FsShared {                               // ROBOT warp: wait for the RNEA warp, then read this lane's column
    const float* col;                                             // fs[i] at col[i * 32]
    __device__ __forceinline__ void operator()(float* fs) const {
        sk_bar_sync(1, 64);
#pragma unroll
        for (int i = 0; i < NU; i++) fs[i] = col[i * 32];
    }
};

template <int NA, int NF> struct RmaskShared {                    // ROBOT warp: the CUBES warp evaluated the pad reach for both roles
    const unsigned* p;
    __device__ __forceinline__ unsigned operator()(const gg::RobotDyn<NA, NF>&) const {
        sk_bar_sync(2, 64);
        return *p;
    }
};

template <int NA, int NF, bool FH, int K>
__global__ void __launch_bounds__(96, 4) stack_prep_kernel(const __grid_constant__ StackArgs A) {
    constexpr int NU = NA + NF;
    typedef gg::StackIn<NA, NF, K> InT;
    typedef Rows<NU, K> R;
    constexpr int WF = sizeof(InT) / sizeof(float);
    constexpr int WFP = WF | 1;                                   // odd row stride: lane = env writes are conflict free
    __shared__ float sm[32 * WFP];
    __shared__ float s_fs[NU * 32];
    __shared__ unsigned s_rmask[32], s_cmask[32];
    __shared__ int s_rit[32];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const int e0 = blockIdx.x * 32 + lane;
    const bool live = e0 < A.n;
    const int e = live ? e0 : A.n - 1;
    const uint32_t gid = (uint32_t)(A.first_env + e);
    const size_t st = (size_t)A.stride;
    float* sp = A.state + e;
    const bool do_reset = stack_reset_decision(A, e, gid);
    const bool active = do_reset ? (A.M.settle_steps > 0) : (A.mode == 0);
    float* rec = sm + lane * WFP;
    InT& In = *reinterpret_cast<InT*>(rec);                       // row of the transposition buffer

    if (role == 1) {
        // ----------------------------------------------------------------------------------------------------------- RNEA
        float q[NU], qd[NU], target[NU], fs[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) {
            const float lq = sp[(R::Q + i) * st], lqd = sp[(R::QD + i) * st];
            q[i] = do_reset ? A.M.home_q[i] : lq; qd[i] = do_reset ? 0.f : lqd;
        }
        const bool use_action = A.mode == 0 && !do_reset;
        // the action row is read ONCE (it may live in pinned host memory: every read crosses PCIe) and, for host-fed steps, kept on the
        // device so that the PD targets persist
        const bool keep = A.action_dev != nullptr && A.mode == 0 && live;
#pragma unroll
        for (int i = 0; i < NU; i++) {
            const float av = (use_action || keep) ? A.action[(size_t)e * NU + i] : 0.f;
            target[i] = use_action ? av : A.M.home_q[i];
            if (keep) A.action_dev[(size_t)e * NU + i] = av;
        }
        gg::robot_rnea<NA, NF, FH>(A.M, A.C, q, qd, target, fs);
#pragma unroll
        for (int i = 0; i < NU; i++) s_fs[i * 32 + lane] = fs[i];
        sk_bar_arrive(1, 64);
    } else if (role == 0) {
        // ---------------------------------------------------------------------------------------------------------- ROBOT
        float q[NU], qd[NU], warm[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) {
            const float lq = sp[(R::Q + i) * st], lqd = sp[(R::QD + i) * st], lw = sp[(R::WARM + i) * st];
            q[i] = do_reset ? A.M.home_q[i] : lq; qd[i] = do_reset ? 0.f : lqd; warm[i] = do_reset ? 0.f : lw;
        }
        int rit = 0;
        FsShared<NU> fsp{s_fs + lane};
        RmaskShared<NA, NF> rmp{s_rmask + lane};
        gg::stack_prep_robot<NA, NF, FH, K>(A.M, A.C, q, qd, warm, A.iterations, A.tol, active && live, In.W, rit, fsp, rmp);
        s_rit[lane] = rit;
    } else {
        // ---------------------------------------------------------------------------------------------------------- CUBES
        gg::ObjState<K> O;
        if (do_reset) spawn_objects<K>(A, e, gid, (uint32_t)A.episode[e], true, O);
        else load_objects<NU, K>(sp, st, O);
        if (do_reset) {
#pragma unroll
            for (int k = 0; k < K; k++)
#pragma unroll
                for (int i = 0; i < 6; i++) In.warm_obj[k][i] = 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < K; k++)
#pragma unroll
                for (int i = 0; i < 6; i++) In.warm_obj[k][i] = sp[(R::WARM + NU + 6 * k + i) * st];
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            float* o = In.obj[k];
            o[0] = O.p[k].x; o[1] = O.p[k].y; o[2] = O.p[k].z; o[3] = O.q[k].w; o[4] = O.q[k].x; o[5] = O.q[k].y; o[6] = O.q[k].z;
            o[7] = O.v[k].x; o[8] = O.v[k].y; o[9] = O.v[k].z; o[10] = O.w[k].x; o[11] = O.w[k].y; o[12] = O.w[k].z;
        }
        unsigned pad_cubes;
        {
            float q[NU];
#pragma unroll
            for (int i = 0; i < NU; i++) { const float lq = sp[(R::Q + i) * st]; q[i] = do_reset ? A.M.home_q[i] : lq; }
            gg::RobotDyn<NA, NF> D;
            gg::robot_pad_frames<NA, NF, FH>(A.M, q, D);
            bool pad_plane;
            gg::stack_pad_reach<NA, NF, K>(A.M, D, O.p, O.q, pad_plane, pad_cubes);
            s_rmask[lane] = ((pad_plane || pad_cubes != 0u) ? 1u : 0u) | (pad_cubes << 1);
            sk_bar_arrive(2, 64);                                 // ROBOT picks the robot's amask bits up after its Cholesky
        }
        s_cmask[lane] = gg::stack_cube_reach<K>(A.M, O, pad_cubes);   // the isolated cubes are solved by stack_cubes_kernel
        if (do_reset && live) {
            // the state of a reset env is rewritten only after every role has read what it needs (joint rows: below, by ROBOT)
            store_objects<NU, K>(sp, st, O);
#pragma unroll
            for (int i = 0; i < 6 * K; i++) sp[(R::WARM + NU + i) * st] = 0.f;
        }
    }
    __syncthreads();
    if (role == 0) {
        if (do_reset && live) {
#pragma unroll
            for (int i = 0; i < NU; i++) { sp[(R::Q + i) * st] = A.M.home_q[i]; sp[(R::QD + i) * st] = 0.f; sp[(R::WARM + i) * st] = 0.f; }
        }
        const unsigned amask = s_rmask[lane] | s_cmask[lane];
        In.W.amask = (float)amask; In.W.iso_iters = (float)s_rit[lane]; In.W.iso_ncon = 0.f;
        if (live) {
            const bool coop = active && amask != 0u;
            A.flags[e] = (do_reset ? 1 : 0) | (active ? 2 : 0) | (coop ? 4 : 0);
            A.amask[e] = (int32_t)amask;
            A.rob_it[e] = (uint8_t)s_rit[lane];
            if (active) {
                // acceleration of the robot block's register solve; the cooperative solve overwrites it when the robot is active
                float* ap = A.acc + e;
#pragma unroll
                for (int i = 0; i < NU; i++) ap[(size_t)i * st] = In.W.arob[i];
                if (coop) {
                    // two lists, drained heavy first (longest-processing-time order shortens the tail of the persistent grid):
                    // heavy = the robot or at least three cubes take part (15+ dofs, pad contacts), light = a pair of cubes
                    const bool heavy = (amask & 1u) || __popc(amask >> 1) >= 3;
                    if (heavy) A.coop_list[atomicAdd(&A.coop_ctr[0], 1)] = e;
                    else A.coop_list[A.stride + atomicAdd(&A.coop_ctr[2], 1)] = e;
                }
            }
        }
    }
    __syncthreads();
    // coalesced env-major copy of the block's records (all three warps)
    const int nlive = min(32, A.n - blockIdx.x * 32);
    float* dst = A.work + (size_t)blockIdx.x * 32 * WF;
    for (int r = role; r < nlive; r += 3)
        for (int i = lane; i < WF; i += 32) dst[(size_t)r * WF + i] = sm[r * WFP + i];
}

// ----------------------------------------------------------------------------------------------------------------- cubes
// One thread per (cube, env): the cubes that can touch nothing but the plane -- the large majority -- are solved here in registers
// (6 x 6 block, <= 4 plane contacts), K times more threads than a thread-per-env pass and none of its serial chain.  Lanes of a warp
// share the cube index and walk consecutive envs, so every state row is one coalesced line.
template <int NU, int K>
__global__ void __launch_bounds__(128) stack_cubes_kernel(const __grid_constant__ StackArgs A) {
    typedef Rows<NU, K> R;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = t / A.stride, e = t - k * A.stride;
    if (k >= K || e >= A.n) return;
    const size_t st = (size_t)A.stride;
    const bool solve = (A.flags[e] & 2) && !((A.amask[e] >> (1 + k)) & 1);
    int nc = 0, it = 0;
    if (solve) {
        const float* c = A.state + e + (R::OBJ + 13 * k) * st;
        const gg::V3 p = gg::mk(c[0], c[st], c[2 * st]);
        const gg::Q4 q = gg::mkq(c[3 * st], c[4 * st], c[5 * st], c[6 * st]);
        const gg::V3 v = gg::mk(c[7 * st], c[8 * st], c[9 * st]);
        const gg::V3 w = gg::mk(c[10 * st], c[11 * st], c[12 * st]);
        float warm[6], a_c[6];
#pragma unroll
        for (int i = 0; i < 6; i++) warm[i] = A.state[e + (R::WARM + NU + 6 * k + i) * st];
        gg::stack_cube_isolated(A.M, p, q, v, w, warm, A.iterations, A.tol, true, a_c, nc, it);
#pragma unroll
        for (int i = 0; i < 6; i++) A.acc[(size_t)(NU + 6 * k + i) * st + e] = a_c[i];
    }
    A.iso_nc[(size_t)k * st + e] = (uint8_t)nc;
    A.iso_it[(size_t)k * st + e] = (uint8_t)it;
}

// ----------------------------------------------------------------------------------------------------------------- solve
#ifndef GG_SOLVE_WARPS
#define GG_SOLVE_WARPS 4
#endif
constexpr int kSolveWarps = GG_SOLVE_WARPS;

// PERSISTENT warps: every warp pops env ids from the list stack_prep_kernel compacted (envs whose bodies can touch each other)
// until it is empty.  A block-per-4-envs launch left 60 % of the warp slots idle -- warps of envs with nothing to solve exit at
// once but their block keeps its 92 KB of shared memory until its slowest warp is done (profiles/r01_v10_stack_*).
template <int NA, int NF, bool FH, int K>
__global__ void __launch_bounds__(32 * kSolveWarps) stack_solve_kernel(const __grid_constant__ StackArgs A) {
    constexpr int NU = NA + NF;
    typedef gg::StackIn<NA, NF, K> InT;
    typedef gg::StackBuf<NU, K> BufT;
    constexpr int WF = sizeof(InT) / sizeof(float);
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* base = smem + (size_t)warp * (sizeof(InT) + sizeof(BufT));
    InT& In = *reinterpret_cast<InT*>(base);
    BufT& B = *reinterpret_cast<BufT*>(base + sizeof(InT));
    const int n_heavy = A.coop_ctr[0];
    const int count = n_heavy + A.coop_ctr[2];                       // heavy list first, then the light one
    WarpDev wc{lane};
#ifdef GG_SOLVE_LOCKSTEP
    __shared__ int s_base;
#endif
    for (;;) {
        int k = 0;
#ifdef GG_SOLVE_LOCKSTEP
        // the warps of a block start their envs together: the solver is one long instruction stream executed about once per env,
        // so warps in the same phase share the instruction cache lines they pull in
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(&A.coop_ctr[1], kSolveWarps);
        __syncthreads();
        if (s_base >= count) break;                               // block-uniform
        k = s_base + warp;
        if (k >= count) continue;
#else
        if (lane == 0) k = atomicAdd(&A.coop_ctr[1], 1);
        k = __shfl_sync(kFullMask, k, 0);
        if (k >= count) break;
#endif
        const int e = k < n_heavy ? A.coop_list[k] : A.coop_list[A.stride + k - n_heavy];
        {
            const float* src = A.work + (size_t)e * WF;
            float* dst = reinterpret_cast<float*>(&In);
            for (int i = lane; i < WF; i += 32) dst[i] = __ldcg(src + i);
        }
        __syncwarp();
        int its = 0, nc = 0;
        gg::stack_solve<NA, NF, FH, K>(A.M, A.T.max_contacts, In, B, A.iterations, A.tol, wc, A.acc + e, A.stride, its, nc);
        if (lane == 0) {
            A.ncon[e] = nc; A.iters[e] = its;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------- finish
template <int NA, int NF, bool FH, int K>
__global__ void __launch_bounds__(128) stack_finish_kernel(const __grid_constant__ StackArgs A) {
    constexpr int NU = NA + NF, NV = NU + 6 * K;
    typedef Rows<NU, K> R;
    const int lane = threadIdx.x & 31;
    const int e0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e0 < A.n;
    const int e = live ? e0 : A.n - 1;
    const uint32_t gid = (uint32_t)(A.first_env + e);
    const size_t st = (size_t)A.stride;
    float* sp = A.state + e;
    const int fl = A.mode == 2 ? 0 : A.flags[e];
    bool do_reset = (fl & 1) != 0;
    const bool active = (fl & 2) != 0;
    float st_steps = 0.f, st_done = 0.f, st_succ = 0.f, st_ret = 0.f, st_len = 0.f, st_nan = 0.f, st_iter = 0.f, st_con = 0.f;

    float q[NU], qd[NU], qa[NV];
    gg::ObjState<K> O;
#pragma unroll
    for (int i = 0; i < NU; i++) { q[i] = sp[(R::Q + i) * st]; qd[i] = sp[(R::QD + i) * st]; }
    load_objects<NU, K>(sp, st, O);
    int steps = A.step_count[e], epi = A.episode[e];
    if (do_reset) { epi += 1; steps = 0; }
    if (active) {
        float a[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) a[i] = A.acc[(size_t)i * st + e];
        gg::integrate_robot<NU>(A.M, q, qd, qa, a);
#pragma unroll
        for (int k = 0; k < K; k++) gg::integrate_cube(A.M, O.p[k], O.q[k], O.v[k], O.w[k], qa + NU + 6 * k, a + NU + 6 * k);
        {
            // contacts / Newton iterations of the step: the cooperative solve (if any) + the register solves of robot and cubes
            int nc = (fl & 4) ? A.ncon[e] : 0, it = (fl & 4) ? A.iters[e] : 0;
            const int ri = A.rob_it[e];
            it = it > ri ? it : ri;
#pragma unroll
            for (int k = 0; k < K; k++) {
                nc += A.iso_nc[(size_t)k * st + e];
                const int ci = A.iso_it[(size_t)k * st + e];
                it = it > ci ? it : ci;
            }
            if (live) { A.ncon[e] = nc; A.iters[e] = it; }
            st_iter = (float)it; st_con = (float)nc;
        }
        float chk = 0.f;
#pragma unroll
        for (int i = 0; i < NU; i++) chk += q[i] + qd[i];
#pragma unroll
        for (int k = 0; k < K; k++)
            chk += O.p[k].x + O.p[k].y + O.p[k].z + O.q[k].w + O.q[k].x + O.q[k].y + O.q[k].z + O.v[k].x + O.v[k].y + O.v[k].z + O.w[k].x +
                   O.w[k].y + O.w[k].z;
        if (!isfinite(chk)) {
            // non-finite env: re-spawn without a substep (counted in the statistics, never thrown from the hot path)
#pragma unroll
            for (int i = 0; i < NU; i++) { q[i] = A.M.home_q[i]; qd[i] = 0.f; }
#pragma unroll
            for (int i = 0; i < NV; i++) qa[i] = 0.f;
            spawn_objects<K>(A, e, gid, (uint32_t)epi, false, O);
            epi += 1; steps = 0; st_nan = 1.f; do_reset = true;
        }
    }
    if (A.mode == 0 && !do_reset) { steps += 1; st_steps = 1.f; }
    float agent[7 + NF], envs[GG_STACK_ENV_STATE_DIM], rew;
    bool succ;
    gg::stack_observe<NA, NF, K>(A.M, A.T, q, O, agent, envs, rew, succ);
    const bool trunc = A.time_limit && steps >= A.M.horizon;
    if (live) {
        const int adim = A.T.agent_qpos ? NU : 7 + NF;
        float* oa = A.obs_agent + (size_t)e * adim;
#pragma unroll
        for (int i = 0; i < 7 + NF; i++) if (i < adim) oa[i] = agent[i];
        float* oe = A.obs_env + (size_t)e * GG_STACK_ENV_STATE_DIM;
#pragma unroll
        for (int i = 0; i < GG_STACK_ENV_STATE_DIM; i++) oe[i] = envs[i];
        if (A.obs_agent_host) {
            float* oh = A.obs_agent_host + (size_t)e * adim;
#pragma unroll
            for (int i = 0; i < 7 + NF; i++) if (i < adim) oh[i] = agent[i];
        }
        if (A.obs_env_host) {
            float* oh = A.obs_env_host + (size_t)e * GG_STACK_ENV_STATE_DIM;
#pragma unroll
            for (int i = 0; i < GG_STACK_ENV_STATE_DIM; i++) oh[i] = envs[i];
        }
        if (A.reward) A.reward[e] = rew;
        if (A.terminated) A.terminated[e] = succ ? 1 : 0;
        if (A.truncated) A.truncated[e] = trunc ? 1 : 0;
        if (A.reward_host) A.reward_host[e] = rew;
        if (A.terminated_host) A.terminated_host[e] = succ ? 1 : 0;
        if (A.truncated_host) A.truncated_host[e] = trunc ? 1 : 0;
        if (A.mode != 2) {
            A.done[e] = (succ || trunc) ? 1 : 0;
            if (active) {
#pragma unroll
                for (int i = 0; i < NU; i++) { sp[(R::Q + i) * st] = q[i]; sp[(R::QD + i) * st] = qd[i]; }
                store_objects<NU, K>(sp, st, O);
#pragma unroll
                for (int i = 0; i < NV; i++) sp[(R::WARM + i) * st] = qa[i];
            }
            A.step_count[e] = steps; A.episode[e] = epi;
            if (A.mode == 0 && (succ || trunc)) { st_done = 1.f; st_succ = succ ? 1.f : 0.f; st_len = (float)steps; }
            if (A.mode == 0) st_ret = rew;
        }
    } else {
        st_steps = st_done = st_succ = st_ret = st_len = st_nan = st_iter = st_con = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { A.coop_ctr[0] = 0; A.coop_ctr[1] = 0; A.coop_ctr[2] = 0; }      // ready for the next step's prep
    if (A.stats != nullptr && A.mode == 0) {
        float v[8] = {st_steps, st_done, st_succ, st_ret, st_len, st_nan, st_iter, st_con};
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFullMask, x, o);
            if (lane == k) mine = x;
        }
        if (lane < 8 && mine != 0.f) atomicAdd(&A.stats[lane], (double)mine);
    }
}

// env-major <-> struct-of-arrays copies (state injection / inspection)
__global__ void stack_state_copy_kernel(float* state, int stride, int n, int NU, int K, float* q, float* qd, float* qacc, int to_soa) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const size_t st = (size_t)stride;
    float* sp = state + e;
    const int NQ = NU + 7 * K, NV = NU + 6 * K, OBJ = 2 * NU, WARM = 2 * NU + 13 * K;
    for (int i = 0; i < NU; i++) {
        if (q) { float* a = q + (size_t)e * NQ + i; float* b = sp + i * st; if (to_soa) *b = *a; else *a = *b; }
        if (qd) { float* a = qd + (size_t)e * NV + i; float* b = sp + (NU + i) * st; if (to_soa) *b = *a; else *a = *b; }
    }
    for (int k = 0; k < K; k++) {
        for (int i = 0; i < 7; i++)
            if (q) { float* a = q + (size_t)e * NQ + NU + 7 * k + i; float* b = sp + (OBJ + 13 * k + i) * st; if (to_soa) *b = *a; else *a = *b; }
        for (int i = 0; i < 6; i++)
            if (qd) { float* a = qd + (size_t)e * NV + NU + 6 * k + i; float* b = sp + (OBJ + 13 * k + 7 + i) * st; if (to_soa) *b = *a; else *a = *b; }
    }
    for (int i = 0; i < NV; i++)
        if (qacc) { float* a = qacc + (size_t)e * NV + i; float* b = sp + (WARM + i) * st; if (to_soa) *b = *a; else *a = *b; }
}

__global__ void stack_stats_kernel(double* stats, double* out, int clear) {
    const int i = threadIdx.x;
    if (i < 8) { out[i] = stats[i]; if (clear) stats[i] = 0.0; }
}

}  // namespace

struct GgStackSim {
    GgModel model;
    GgStackTask task;
    GgConfig cfg;
    int n, device, stride, nu, nv, kind;                     // kind: 0 = <7,2,slide> Panda, 1 = <5,1,hinge> SO-101
    float* state; int32_t* step_count; int32_t* episode;     // caller-owned
    uint8_t* done; uint8_t* flags; double* stats;            // library-owned
    float* work; float* acc; int32_t* ncon; int32_t* iters;
    int32_t* coop_list; int32_t* coop_ctr;
    int32_t* amask; uint8_t* rob_it; uint8_t* iso_nc; uint8_t* iso_it;
    int32_t* slot;                                            // allocated on first gg_stack_export_successful
    uint32_t step_index;
    GgPinCache pins;                                          // address ranges already verified as pinned host memory
    size_t solve_smem;
    int solve_blocks;                                         // persistent grid of the cooperative solver
    cudaStream_t side;                                        // the cubes kernel runs beside the cooperative solve (fork / join per step)
    cudaEvent_t ev_fork, ev_join;
};

typedef gg::StackIn<7, 2, 5> FrankaIn;
typedef gg::StackBuf<9, 5> FrankaBuf;
typedef gg::StackIn<5, 1, 5> So101In;
typedef gg::StackBuf<6, 5> So101Buf;

template <int NA, int NF, bool FH, int K>
static void stack_launch_t(GgStackSim* sim, const StackArgs& A, cudaStream_t s);

static int stack_launch(GgStackSim* sim, StackArgs& A, gg_stream_t stream) {
    if (sim->state == nullptr) return gg_internal_fail("gg_stack_bind_state has not been called%s", "");
    A.M = sim->model; gg::make_chain(A.M, A.C); A.T = sim->task; A.n = sim->n; A.stride = sim->stride;
    A.state = sim->state; A.step_count = sim->step_count; A.episode = sim->episode; A.done = sim->done;
    A.iterations = sim->cfg.solver_iterations; A.tol = sim->cfg.solver_tolerance;
    A.auto_reset = sim->cfg.auto_reset; A.time_limit = sim->cfg.time_limit; A.reset_prob = sim->cfg.reset_prob;
    A.seed = sim->cfg.seed; A.first_env = sim->cfg.first_env_id; A.step_index = sim->step_index;
    A.stats = sim->stats; A.work = sim->work; A.acc = sim->acc; A.ncon = sim->ncon; A.iters = sim->iters; A.flags = sim->flags;
    A.coop_list = sim->coop_list; A.coop_ctr = sim->coop_ctr;
    A.amask = sim->amask; A.rob_it = sim->rob_it; A.iso_nc = sim->iso_nc; A.iso_it = sim->iso_it;
    cudaStream_t s = (cudaStream_t)stream;
    GgDeviceGuard dev_guard(sim->device); SK_CUDA(dev_guard.err);
    if (sim->kind == 0) stack_launch_t<7, 2, false, 5>(sim, A, s);
    else stack_launch_t<5, 1, true, 5>(sim, A, s);
    gg_internal_count_launches(A.mode != 2 ? 4 : 1);
    SK_CUDA(cudaGetLastError());
    return 0;
}

template <int NA, int NF, bool FH, int K>
static void stack_launch_t(GgStackSim* sim, const StackArgs& A, cudaStream_t s) {
    if (A.mode != 2) {
        stack_prep_kernel<NA, NF, FH, K><<<(sim->n + 31) / 32, 96, 0, s>>>(A);
        // the register solves of the isolated cubes and the cooperative solve touch disjoint dofs: fork the former onto a side
        // stream so that they share the SMs (the solver is shared-memory bound at 8 warps / SM and leaves registers and warp slots
        // free), join before the finish kernel.  Event fork / join is capturable, so the step still records into a CUDA graph.
        cudaEventRecord(sim->ev_fork, s);
        cudaStreamWaitEvent(sim->side, sim->ev_fork, 0);
        stack_cubes_kernel<NA + NF, K><<<(K * sim->stride + 127) / 128, 128, 0, sim->side>>>(A);
        cudaEventRecord(sim->ev_join, sim->side);
        const int need = (sim->n + kSolveWarps - 1) / kSolveWarps;
        stack_solve_kernel<NA, NF, FH, K><<<need < sim->solve_blocks ? need : sim->solve_blocks, 32 * kSolveWarps, sim->solve_smem, s>>>(A);
        cudaStreamWaitEvent(s, sim->ev_join, 0);
    }
    stack_finish_kernel<NA, NF, FH, K><<<(sim->n + 127) / 128, 128, 0, s>>>(A);
}

extern "C" {

int gg_stack_create(const GgModel* model, const GgStackTask* task, int n_envs, int device, GgStackSim** out) {
    if (!model || !task || !out) return gg_internal_fail("gg_stack_create: null argument%s", "");
    if (n_envs <= 0) return gg_internal_fail("gg_stack_create: n_envs must be positive%s", "");
    int kind = -1;
    if (model->n_arm == 7 && model->n_fing == 2 && !model->fing_hinge) kind = 0;
    else if (model->n_arm == 5 && model->n_fing == 1 && model->fing_hinge) kind = 1;
    if (kind < 0) return gg_internal_fail("gg_stack_create: unsupported articulation (need 7 hinges + 2 slides or 5 hinges + 1 hinge)%s", "");
    if (task->n_obj != 5) return gg_internal_fail("gg_stack_create: the stacking scene has 5 cubes (cube_1, cube_2, 3 distractors)%s", "");
    if (task->max_contacts < 1 || task->max_contacts > GG_STACK_MAX_CONTACTS) return gg_internal_fail("gg_stack_create: max_contacts out of range%s", "");
    if (model->n_pads < 0 || model->n_pads > GG_MAX_PADS) return gg_internal_fail("gg_stack_create: too many pads%s", "");
    if (model->power != 2.0f) return gg_internal_fail("gg_stack_create: solimp power must be 2%s", "");
    if (model->settle_steps < 0 || model->settle_steps > 1) return gg_internal_fail("gg_stack_create: settle_steps must be 0 or 1%s", "");
    int ndev = 0;
    SK_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return gg_internal_fail("gg_stack_create: bad device index%s", "");
    GgDeviceGuard dev_guard(device); SK_CUDA(dev_guard.err);
    GgStackSim* s = new (std::nothrow) GgStackSim();
    if (!s) return gg_internal_fail("gg_stack_create: out of host memory%s", "");
    memset(s, 0, sizeof(*s));
    s->model = *model; s->task = *task; s->n = n_envs; s->device = device; s->kind = kind;
    s->nu = model->n_arm + model->n_fing; s->nv = s->nu + 6 * task->n_obj;
    s->stride = (n_envs + 31) / 32 * 32;
    s->cfg.solver_iterations = 10; s->cfg.solver_tolerance = 1e-5f;
    s->solve_smem = kSolveWarps * (kind == 0 ? sizeof(FrankaIn) + sizeof(FrankaBuf) : sizeof(So101In) + sizeof(So101Buf));
    cudaError_t e[13];
    e[0] = cudaMalloc(&s->done, (size_t)s->stride);
    e[1] = cudaMalloc(&s->flags, (size_t)s->stride);
    e[2] = cudaMalloc(&s->stats, 8 * sizeof(double));
    e[3] = cudaMalloc(&s->work, (size_t)s->stride * (kind == 0 ? sizeof(FrankaIn) : sizeof(So101In)));
    e[4] = cudaMalloc(&s->acc, (size_t)s->stride * s->nv * sizeof(float));
    e[5] = cudaMalloc(&s->ncon, (size_t)s->stride * sizeof(int32_t));
    e[6] = cudaMalloc(&s->iters, (size_t)s->stride * sizeof(int32_t));
    e[7] = cudaMalloc(&s->coop_list, 2 * (size_t)s->stride * sizeof(int32_t));
    e[8] = cudaMalloc(&s->coop_ctr, 4 * sizeof(int32_t));
    e[9] = cudaMalloc(&s->amask, (size_t)s->stride * sizeof(int32_t));
    e[10] = cudaMalloc(&s->rob_it, (size_t)s->stride);
    e[11] = cudaMalloc(&s->iso_nc, (size_t)s->stride * task->n_obj);
    e[12] = cudaMalloc(&s->iso_it, (size_t)s->stride * task->n_obj);
    bool ok = true;
    for (int i = 0; i < 13; i++) ok = ok && e[i] == cudaSuccess;
    {
        // persistent solver grid: as many blocks as fit at once (shared-memory bound), one wave on every SM
        cudaDeviceProp prop;
        int per_sm = 0;
        if (ok && cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
            per_sm = (int)(prop.sharedMemPerMultiprocessor / (s->solve_smem + 1024));
            if (per_sm < 1) per_sm = 1;
            s->solve_blocks = prop.multiProcessorCount * per_sm;
        } else {
            s->solve_blocks = 148 * 2;
        }
    }
    if (ok && kind == 0) ok = cudaFuncSetAttribute(stack_solve_kernel<7, 2, false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->solve_smem) == cudaSuccess;
    if (ok && kind == 1) ok = cudaFuncSetAttribute(stack_solve_kernel<5, 1, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->solve_smem) == cudaSuccess;
    if (!ok) {
        cudaFree(s->done); cudaFree(s->flags); cudaFree(s->stats); cudaFree(s->work); cudaFree(s->acc); cudaFree(s->ncon); cudaFree(s->iters);
        cudaFree(s->coop_list); cudaFree(s->coop_ctr); cudaFree(s->amask); cudaFree(s->rob_it); cudaFree(s->iso_nc); cudaFree(s->iso_it);
        delete s;
        return gg_internal_fail("gg_stack_create: device allocation failed%s", "");
    }
    cudaMemset(s->done, 0, (size_t)s->stride);
    cudaMemset(s->flags, 0, (size_t)s->stride);
    cudaMemset(s->stats, 0, 8 * sizeof(double));
    cudaMemset(s->ncon, 0, (size_t)s->stride * sizeof(int32_t));
    cudaMemset(s->iters, 0, (size_t)s->stride * sizeof(int32_t));
    ok = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { gg_stack_destroy(s); return gg_internal_fail("gg_stack_create: stream / event creation failed%s", ""); }
    cudaMemset(s->coop_ctr, 0, 4 * sizeof(int32_t));
    cudaMemset(s->amask, 0, (size_t)s->stride * sizeof(int32_t));
    cudaMemset(s->rob_it, 0, (size_t)s->stride);
    cudaMemset(s->iso_nc, 0, (size_t)s->stride * task->n_obj);
    cudaMemset(s->iso_it, 0, (size_t)s->stride * task->n_obj);
    *out = s;
    return 0;
}

int gg_stack_destroy(GgStackSim* s) {
    if (!s) return 0;
    GgDeviceGuard dev_guard(s->device);
    if (s->side) { cudaStreamSynchronize(s->side); cudaStreamDestroy(s->side); }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    cudaFree(s->done); cudaFree(s->flags); cudaFree(s->stats); cudaFree(s->work); cudaFree(s->acc); cudaFree(s->ncon); cudaFree(s->iters);
    cudaFree(s->coop_list); cudaFree(s->coop_ctr); cudaFree(s->slot);
    cudaFree(s->amask); cudaFree(s->rob_it); cudaFree(s->iso_nc); cudaFree(s->iso_it);
    delete s;
    return 0;
}

int gg_stack_configure(GgStackSim* sim, const GgConfig* cfg) {
    if (!sim || !cfg) return gg_internal_fail("gg_stack_configure: null argument%s", "");
    // <= 255: the per-env iteration counters of the register solves (rob_it, iso_it) are bytes
    if (cfg->solver_iterations < 1 || cfg->solver_iterations > 255) return gg_internal_fail("gg_stack_configure: solver_iterations out of range [1, 255]%s", "");
    if (!(cfg->reset_prob >= 0.f && cfg->reset_prob <= 1.f)) return gg_internal_fail("gg_stack_configure: reset_prob out of range%s", "");
    sim->cfg = *cfg;
    return 0;
}

int gg_stack_agent_dim(const GgStackSim* sim) { return sim ? (sim->task.agent_qpos ? sim->nu : 7 + sim->model.n_fing) : -1; }
int gg_stack_state_rows(const GgStackSim* sim) { return sim ? 3 * sim->nu + 19 * sim->task.n_obj : -1; }
int gg_stack_state_stride(const GgStackSim* sim) { return sim ? sim->stride : -1; }

int gg_stack_bind_state(GgStackSim* sim, float* state, int32_t* step_count, int32_t* episode) {
    if (!sim || !state || !step_count || !episode) return gg_internal_fail("gg_stack_bind_state: null argument%s", "");
    sim->state = state; sim->step_count = step_count; sim->episode = episode;
    return 0;
}

int gg_stack_step(GgStackSim* sim, const float* action, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated,
                  uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream) {
    if (!sim || !action || !obs_agent || !obs_env) return gg_internal_fail("gg_stack_step: null argument%s", "");
    StackArgs A;
    memset(&A, 0, sizeof(A));
    A.action = action; A.obs_agent = obs_agent; A.obs_env = obs_env; A.reward = reward; A.terminated = terminated;
    A.truncated = truncated; A.reset_mask = reset_mask; A.mode = 0;
    const int rc = stack_launch(sim, A, stream);
    sim->step_index++;
    return rc;
}

int gg_stack_step_host(GgStackSim* sim, const GgHostIO* io, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated,
                       uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream) {
    if (!sim || !io || !io->action || !obs_agent || !obs_env) return gg_internal_fail("gg_stack_step_host: null argument%s", "");
    const void* ptrs[6] = {io->action, io->reward, io->terminated, io->truncated, io->obs_agent, io->obs_env};
    for (int k = 0; k < 6; k++)
        if (ptrs[k] && !gg_internal_pinned(&sim->pins, ptrs[k]))
            return gg_internal_fail("gg_stack_step_host: action / mirrors must be page-locked (pinned) host memory%s", "");
    StackArgs A;
    memset(&A, 0, sizeof(A));
    A.action = io->action; A.obs_agent = obs_agent; A.obs_env = obs_env; A.reward = reward; A.terminated = terminated;
    A.truncated = truncated; A.reset_mask = reset_mask; A.mode = 0;
    A.reward_host = io->reward; A.terminated_host = io->terminated; A.truncated_host = io->truncated;
    A.obs_agent_host = io->obs_agent; A.obs_env_host = io->obs_env; A.action_dev = io->action_dev;
    const int rc = stack_launch(sim, A, stream);
    sim->step_index++;
    if (rc != 0) return rc;
    if (io->wait) SK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int gg_stack_reset(GgStackSim* sim, const uint8_t* mask, const float* obj_xy, float* obs_agent, float* obs_env, gg_stream_t stream) {
    if (!sim || !obs_agent || !obs_env) return gg_internal_fail("gg_stack_reset: null argument%s", "");
    StackArgs A;
    memset(&A, 0, sizeof(A));
    A.obs_agent = obs_agent; A.obs_env = obs_env; A.reset_mask = mask; A.obj_xy = obj_xy; A.mode = 1;
    return stack_launch(sim, A, stream);
}

int gg_stack_observe(GgStackSim* sim, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated, gg_stream_t stream) {
    if (!sim || !obs_agent || !obs_env) return gg_internal_fail("gg_stack_observe: null argument%s", "");
    StackArgs A;
    memset(&A, 0, sizeof(A));
    A.obs_agent = obs_agent; A.obs_env = obs_env; A.reward = reward; A.terminated = terminated; A.mode = 2;
    return stack_launch(sim, A, stream);
}

static int stack_state_copy(GgStackSim* sim, float* q, float* qd, float* qacc, int to_soa, gg_stream_t stream) {
    if (!sim || sim->state == nullptr) return gg_internal_fail("stack state copy: state not bound%s", "");
    GgDeviceGuard dev_guard(sim->device); SK_CUDA(dev_guard.err);
    stack_state_copy_kernel<<<(sim->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sim->state, sim->stride, sim->n, sim->nu, sim->task.n_obj, q, qd,
                                                                                   qacc, to_soa);
    gg_internal_count_launches(1);
    SK_CUDA(cudaGetLastError());
    return 0;
}
int gg_stack_get_state(GgStackSim* sim, float* q, float* qd, float* qacc, gg_stream_t stream) { return stack_state_copy(sim, q, qd, qacc, 0, stream); }
int gg_stack_set_state(GgStackSim* sim, const float* q, const float* qd, const float* qacc, gg_stream_t stream) {
    return stack_state_copy(sim, (float*)q, (float*)qd, (float*)qacc, 1, stream);
}

int gg_stack_stats(GgStackSim* sim, double* out8, int clear, gg_stream_t stream) {
    if (!sim || !out8) return gg_internal_fail("gg_stack_stats: null argument%s", "");
    GgDeviceGuard dev_guard(sim->device); SK_CUDA(dev_guard.err);
    stack_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sim->stats, out8, clear);
    gg_internal_count_launches(1);
    SK_CUDA(cudaGetLastError());
    return 0;
}

int gg_stack_frame_dim(const GgStackSim* sim) { return sim ? gg_stack_agent_dim(sim) + GG_STACK_ENV_STATE_DIM + sim->nu + 1 : -1; }

int gg_stack_record(GgStackSim* sim, float* traj, int32_t* ever, int t, int horizon, const float* action, const float* obs_agent,
                    const float* obs_env, const float* reward, gg_stream_t stream) {
    if (!sim) return gg_internal_fail("gg_stack_record: null argument%s", "");
    return gg_internal_record(sim->device, sim->n, gg_stack_agent_dim(sim), GG_STACK_ENV_STATE_DIM, sim->nu, traj, ever, t, horizon, action,
                              obs_agent, obs_env, reward, stream);
}

int gg_stack_export_successful(GgStackSim* sim, const float* traj, const int32_t* ever, int horizon, float* out, int32_t* out_env_ids,
                               int32_t* out_count, gg_stream_t stream) {
    if (!sim) return gg_internal_fail("gg_stack_export_successful: null argument%s", "");
    return gg_internal_export(sim->device, sim->n, sim->stride, gg_stack_frame_dim(sim), &sim->slot, traj, ever, horizon, out, out_env_ids,
                              out_count, stream);
}

int gg_stack_info(GgStackSim* sim, int32_t* ncon, int32_t* iters, gg_stream_t stream) {
    if (!sim) return gg_internal_fail("gg_stack_info: null argument%s", "");
    GgDeviceGuard dev_guard(sim->device); SK_CUDA(dev_guard.err);
    if (ncon) SK_CUDA(cudaMemcpyAsync(ncon, sim->ncon, (size_t)sim->n * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (iters) SK_CUDA(cudaMemcpyAsync(iters, sim->iters, (size_t)sim->n * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

}  // extern "C"
