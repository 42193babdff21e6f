// CPU BASELINE ONLY (bench.py's cpu_baseline / --impl reference legs): the SPECIALISED per-env pipeline -- the same header the CUDA
// kernels are built from (gym_genesis_b200/csrc/gg_env.cuh: unrolled 9-dof chain, register Cholesky, block-diagonal solvers) --
// compiled for the host with -O3 -march=native and run on all host threads.  It is the honest "good CPU implementation" to quote
// beside the GPU number: the generic oracle (oracle.cpp: arbitrary trees, dense Jacobians, dense Newton)
// is 3-4x slower by construction.  Test / benchmark infrastructure: the product never loads it.
//   whole workload step in native code: masked reset (spawn from caller-supplied xy) + settle rule + substep + observation + flags
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>
#include "gg_env.cuh"

namespace {
template <int NA, int NF, bool FH>
void run_range(const GgModel* M, const gg::GgChain* C, int lo, int hi, float* q, float* qd, float* qacc, const float* target,
               const uint8_t* reset_mask, const float* xy, int iterations, float tol, float* agent, float* envs, float* reward, uint8_t* term) {
    constexpr int NU = NA + NF, NQ = NU + 7, NV = NU + 6;
    for (int e = lo; e < hi; e++) {
        gg::EnvState<NU> s;
        float* pq = q + (size_t)e * NQ; float* pv = qd + (size_t)e * NV; float* pa = qacc + (size_t)e * NV;
        float tgt[NU];
        const bool rs = reset_mask && reset_mask[e];
        if (rs) {
            gg::env_reset<NU>(*M, s, xy[2 * e], xy[2 * e + 1]);
            for (int i = 0; i < NU; i++) tgt[i] = M->home_q[i];
        } else {
            for (int i = 0; i < NU; i++) { s.q[i] = pq[i]; s.qd[i] = pv[i]; tgt[i] = target[(size_t)e * NU + i]; }
            s.cp = gg::mk(pq[NU], pq[NU + 1], pq[NU + 2]); s.cq = gg::mkq(pq[NU + 3], pq[NU + 4], pq[NU + 5], pq[NU + 6]);
            s.cv = gg::mk(pv[NU], pv[NU + 1], pv[NU + 2]); s.cw = gg::mk(pv[NU + 3], pv[NU + 4], pv[NU + 5]);
            for (int i = 0; i < NV; i++) s.qacc[i] = pa[i];
        }
        gg::StepInfo si;
        if (!rs || M->settle_steps > 0) gg::env_substep<NA, NF, FH>(*M, *C, s, tgt, iterations, tol, si);
        for (int i = 0; i < NU; i++) { pq[i] = s.q[i]; pv[i] = s.qd[i]; }
        pq[NU] = s.cp.x; pq[NU + 1] = s.cp.y; pq[NU + 2] = s.cp.z;
        pq[NU + 3] = s.cq.w; pq[NU + 4] = s.cq.x; pq[NU + 5] = s.cq.y; pq[NU + 6] = s.cq.z;
        pv[NU] = s.cv.x; pv[NU + 1] = s.cv.y; pv[NU + 2] = s.cv.z; pv[NU + 3] = s.cw.x; pv[NU + 4] = s.cw.y; pv[NU + 5] = s.cw.z;
        for (int i = 0; i < NV; i++) pa[i] = s.qacc[i];
        bool ok;
        gg::env_observe<NA, NF>(*M, s, agent + (size_t)e * (7 + NF), envs + (size_t)e * 11, reward[e], ok);
        term[e] = ok;
    }
}
}  // namespace

extern "C" int hp_sizeof_model(void) { return (int)sizeof(GgModel); }

// The whole benchmark workload in native code: `steps` workload steps of n envs, every thread owns a contiguous range of envs for the
// whole run (envs are independent: no synchronisation at all between steps).  Per step and env: Bernoulli(reset_prob) masked reset
// (uniform spawn in the model's range, home pose, settle rule) or one substep with the step's action row (pool of `npool` action
// batches, cycled), then the observation / reward / flag evaluation.  RNG: one 64-bit LCG stream per env (the draws only have to be
// statistically the workload's; nothing is compared against them).
extern "C" void hp_run_workload(const GgModel* M, int n, int steps, int nthreads, float* q, float* qd, float* qacc, const float* pool, int npool,
                                float reset_prob, uint64_t seed, int iterations, float tol, float* agent, float* envs, float* reward, uint8_t* term) {
    gg::GgChain C;
    gg::make_chain(*M, C);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n;
    const int nu = M->n_arm + M->n_fing;
    auto work = [&](int lo, int hi) {
        std::vector<uint8_t> mask(hi - lo);
        std::vector<float> xy(2 * (size_t)(hi - lo));
        std::vector<uint64_t> rng(hi - lo);
        for (int e = lo; e < hi; e++) rng[e - lo] = seed * 0x9E3779B97F4A7C15ull + (uint64_t)e * 0xD1B54A32D192ED03ull + 1;
        auto u01 = [](uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 40) * (1.0 / 16777216.0)); };
        for (int k = 0; k < steps; k++) {
            const float* act = pool + (size_t)(k % npool) * n * nu;
            for (int e = lo; e < hi; e++) {
                uint64_t& s = rng[e - lo];
                mask[e - lo] = u01(s) < reset_prob;
                if (mask[e - lo]) {
                    xy[2 * (e - lo)] = M->spawn_lo[0] + u01(s) * (M->spawn_hi[0] - M->spawn_lo[0]);
                    xy[2 * (e - lo) + 1] = M->spawn_lo[1] + u01(s) * (M->spawn_hi[1] - M->spawn_lo[1]);
                }
            }
            // run_range indexes the per-env arrays by absolute env id; mask / xy are relative to `lo`
            if (M->n_arm == 7) run_range<7, 2, false>(M, &C, lo, hi, q, qd, qacc, act, mask.data() - lo, xy.data() - 2 * (size_t)lo, iterations, tol, agent, envs, reward, term);
            else run_range<5, 1, true>(M, &C, lo, hi, q, qd, qacc, act, mask.data() - lo, xy.data() - 2 * (size_t)lo, iterations, tol, agent, envs, reward, term);
        }
    };
    if (nthreads == 1) { work(0, n); return; }
    std::vector<std::thread> th;
    const int chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        const int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
}

extern "C" void hp_step(const GgModel* M, int n, float* q, float* qd, float* qacc, const float* target, const uint8_t* reset_mask,
                        const float* xy, int iterations, float tol, float* agent, float* envs, float* reward, uint8_t* term, int nthreads) {
    gg::GgChain C;
    gg::make_chain(*M, C);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n;
    auto work = [&](int lo, int hi) {
        if (M->n_arm == 7) run_range<7, 2, false>(M, &C, lo, hi, q, qd, qacc, target, reset_mask, xy, iterations, tol, agent, envs, reward, term);
        else run_range<5, 1, true>(M, &C, lo, hi, q, qd, qacc, target, reset_mask, xy, iterations, tol, agent, envs, reward, term);
    };
    if (nthreads == 1) { work(0, n); return; }
    std::vector<std::thread> th;
    const int chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        const int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
}
