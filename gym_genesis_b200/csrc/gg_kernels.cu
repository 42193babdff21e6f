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

#include <atomic>
#include <new>

#include "gg_b200.h"
#include "gg_env.cuh"

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
    GgModel M;
    int n, stride;
    float* state; int32_t* step_count; int32_t* episode; uint8_t* done;
    const float* action; float* obs_agent; float* obs_env; float* reward;
    uint8_t* terminated; uint8_t* truncated;
    const uint8_t* reset_mask; const float* cube_xy;
    int mode;                 // 0 = env.step, 1 = reset (+ settle) of masked envs, 2 = observe only
    int iterations; float tol;
    int auto_reset, time_limit; float reset_prob;
    uint64_t seed; int64_t first_env; uint32_t step_index;
    double* stats;
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

template <int NU> __device__ __forceinline__ bool state_finite(const gg::EnvState<NU>& s) {
    float acc = s.cp.x + s.cp.y + s.cp.z + s.cq.w + s.cq.x + s.cq.y + s.cq.z + s.cv.x + s.cv.y + s.cv.z + s.cw.x + s.cw.y + s.cw.z;
#pragma unroll
    for (int i = 0; i < NU; i++) acc += s.q[i] + s.qd[i];
    return isfinite(acc);
}

constexpr int kBlock = 128;      // default threads per block
constexpr int kMaxBlock = 1024;

template <int NA, int NF, bool FH>
__global__ void __launch_bounds__(256) step_kernel(const __grid_constant__ StepArgs A) {
    constexpr int NU = NA + NF;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e < A.n;
    float st_steps = 0.f, st_done = 0.f, st_succ = 0.f, st_ret = 0.f, st_len = 0.f, st_nan = 0.f, st_iter = 0.f, st_con = 0.f;
    if (live) {
        gg::EnvState<NU> s;
        load_state<NU>(A.state, A.stride, e, s);
        int steps = A.step_count[e], epi = A.episode[e];
        const uint32_t gid = (uint32_t)(A.first_env + e);
        bool do_reset = false;
        if (A.mode == 0) {
            do_reset = (A.reset_mask != nullptr && A.reset_mask[e] != 0) || (A.auto_reset && A.done[e] != 0);
            if (A.reset_prob > 0.f) {
                uint32_t o[4];
                gg::philox4x32_10(gid, A.step_index, 0x52535421u, 0u, (uint32_t)A.seed, (uint32_t)(A.seed >> 32), o);
                do_reset = do_reset || gg::u01(o[0]) < A.reset_prob;
            }
        } else if (A.mode == 1) {
            do_reset = A.reset_mask == nullptr || A.reset_mask[e] != 0;
        }
        float tgt[NU];
        int nsub = 0;
        if (do_reset) {
            float x, y;
            if (A.cube_xy != nullptr) { x = A.cube_xy[2 * e]; y = A.cube_xy[2 * e + 1]; }
            else gg::spawn_draw(A.M, A.seed, gid, (uint32_t)epi, x, y);
            gg::env_reset<NU>(A.M, s, x, y);
            epi += 1; steps = 0;
#pragma unroll
            for (int i = 0; i < NU; i++) tgt[i] = A.M.home_q[i];
            nsub = A.M.settle_steps;
        } else if (A.mode == 0) {
#pragma unroll
            for (int i = 0; i < NU; i++) tgt[i] = A.action[(size_t)e * NU + i];
            nsub = 1;
        }
        for (int k = 0; k < nsub; k++) {
            gg::StepInfo info;
            gg::env_substep<NA, NF, FH>(A.M, s, tgt, A.iterations, A.tol, info);
            st_iter += (float)info.iters; st_con += (float)info.ncon;
        }
        if (nsub > 0 && !state_finite<NU>(s)) {       // diverged env: flag, re-spawn, keep the batch alive
            float x, y;
            gg::spawn_draw(A.M, A.seed, gid, (uint32_t)epi, x, y);
            gg::env_reset<NU>(A.M, s, x, y);
            epi += 1; steps = 0; st_nan = 1.f;
            do_reset = true;
        }
        if (A.mode == 0 && !do_reset) { steps += 1; st_steps = 1.f; }
        float agent[7 + NF], envs[GG_ENV_STATE_DIM], rew; bool succ;
        gg::env_observe<NA, NF>(A.M, s, agent, envs, rew, succ);
        const bool trunc = A.time_limit && steps >= A.M.horizon;
#pragma unroll
        for (int i = 0; i < 7 + NF; i++) A.obs_agent[(size_t)e * (7 + NF) + i] = agent[i];
#pragma unroll
        for (int i = 0; i < GG_ENV_STATE_DIM; i++) A.obs_env[(size_t)e * GG_ENV_STATE_DIM + i] = envs[i];
        if (A.reward) A.reward[e] = rew;
        if (A.terminated) A.terminated[e] = succ ? 1 : 0;
        if (A.truncated) A.truncated[e] = trunc ? 1 : 0;
        if (A.mode != 2) {
            A.done[e] = (succ || trunc) ? 1 : 0;
            if (nsub > 0 || do_reset) store_state<NU>(A.state, A.stride, e, s);
            A.step_count[e] = steps; A.episode[e] = epi;
            if (A.mode == 0 && (succ || trunc)) { st_done = 1.f; st_succ = succ ? 1.f : 0.f; st_len = (float)steps; }
            if (A.mode == 0) st_ret = rew;
        }
    }
    if (A.stats != nullptr && A.mode == 0) {
        float v[8] = {st_steps, st_done, st_succ, st_ret, st_len, st_nan, st_iter, st_con};
        __shared__ float red[8][kMaxBlock / 32];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = x;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            float x = 0.f;
#pragma unroll
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) x += red[threadIdx.x][w];
            if (x != 0.f) atomicAdd(&A.stats[threadIdx.x], (double)x);
        }
    }
}

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

__global__ void stats_kernel(double* stats, double* out, int clear) {
    int i = threadIdx.x;
    if (i < 8) { out[i] = stats[i]; if (clear) stats[i] = 0.0; }
}

}  // namespace

struct GgSim {
    GgModel model;
    GgConfig cfg;
    int n, device, stride, kind;        // kind: 0 = <7,2,slide>, 1 = <5,1,hinge>
    int block;                          // threads per block of step_kernel
    float* state; int32_t* step_count; int32_t* episode;
    uint8_t* done; double* stats;       // library-owned (allocated once in gg_create)
    uint32_t step_index;
};

static int launch_step(GgSim* sim, StepArgs& A, gg_stream_t stream) {
    if (sim->state == nullptr) return fail("gg_bind_state has not been called%s");
    A.M = sim->model; A.n = sim->n; A.stride = sim->stride;
    A.state = sim->state; A.step_count = sim->step_count; A.episode = sim->episode; A.done = sim->done;
    A.iterations = sim->cfg.solver_iterations; A.tol = sim->cfg.solver_tolerance;
    A.auto_reset = sim->cfg.auto_reset; A.time_limit = sim->cfg.time_limit; A.reset_prob = sim->cfg.reset_prob;
    A.seed = sim->cfg.seed; A.first_env = sim->cfg.first_env_id; A.step_index = sim->step_index;
    A.stats = sim->stats;
    const int grid = (sim->n + sim->block - 1) / sim->block;
    GG_CUDA(cudaSetDevice(sim->device));
    if (sim->kind == 0) step_kernel<7, 2, false><<<grid, sim->block, 0, (cudaStream_t)stream>>>(A);
    else step_kernel<5, 1, true><<<grid, sim->block, 0, (cudaStream_t)stream>>>(A);
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
    GG_CUDA(cudaSetDevice(device));
    GgSim* s = new (std::nothrow) GgSim();
    if (!s) return fail("gg_create: out of host memory%s");
    memset(s, 0, sizeof(*s));
    s->model = *model; s->n = n_envs; s->device = device; s->kind = kind;
    s->stride = (n_envs + 31) / 32 * 32;
    s->block = kBlock;
    if (const char* eb = getenv("GG_BLOCK")) { int b = atoi(eb); if (b >= 32 && b <= 256 && b % 32 == 0) s->block = b; }
    s->cfg.solver_iterations = 10; s->cfg.solver_tolerance = 1e-5f;
    cudaError_t e1 = cudaMalloc(&s->done, (size_t)s->stride);
    cudaError_t e2 = cudaMalloc(&s->stats, 8 * sizeof(double));
    if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaFree(s->done); cudaFree(s->stats); delete s; return fail("gg_create: cudaMalloc failed%s"); }
    cudaMemset(s->done, 0, (size_t)s->stride);
    cudaMemset(s->stats, 0, 8 * sizeof(double));
    *out = s;
    return 0;
}

int gg_destroy(GgSim* sim) {
    if (!sim) return 0;
    cudaSetDevice(sim->device);
    cudaFree(sim->done); cudaFree(sim->stats);
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
    int rc = launch_step(sim, A, stream);
    sim->step_index++;
    return rc;
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
    GG_CUDA(cudaSetDevice(sim->device));
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

int gg_stats(GgSim* sim, double* out8, int clear, gg_stream_t stream) {
    if (!sim || !out8) return fail("gg_stats: null argument%s");
    GG_CUDA(cudaSetDevice(sim->device));
    stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sim->stats, out8, clear);
    g_launches.fetch_add(1);
    GG_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
