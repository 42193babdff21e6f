// TEST HARNESS ONLY: runs the multi-object step pipeline of gym_genesis_b200/csrc/gg_stack.cuh on the host so that the
// kernel math -- including the WARP-COOPERATIVE solve, emulated with 32 host threads per warp and a barrier behind every
// warp collective -- can be checked against the oracle in the CPU-only test tier.  The product never loads this library.
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#include "gg_stack.cuh"

namespace {

struct HostWarp {
    pthread_barrier_t bar;
    float f[32];
    int i[32];
    HostWarp() { pthread_barrier_init(&bar, nullptr, 32); }
    ~HostWarp() { pthread_barrier_destroy(&bar); }
};

struct WarpHost {
    int lane;
    HostWarp* w;
    void sync() { pthread_barrier_wait(&w->bar); }
    float shfl(float v, int src) { w->f[lane] = v; sync(); float r = w->f[src]; sync(); return r; }
    int shfli(int v, int src) { w->i[lane] = v; sync(); int r = w->i[src]; sync(); return r; }
    float sum(float v) { w->f[lane] = v; sync(); float r = 0.f; for (int l = 0; l < 32; l++) r += w->f[l]; sync(); return r; }
    float max(float v) { w->f[lane] = v; sync(); float r = w->f[0]; for (int l = 1; l < 32; l++) r = fmaxf(r, w->f[l]); sync(); return r; }
    bool any(bool p) { w->i[lane] = p ? 1 : 0; sync(); int r = 0; for (int l = 0; l < 32; l++) r |= w->i[l]; sync(); return r != 0; }
    unsigned ballot(bool p) { w->i[lane] = p ? 1 : 0; sync(); unsigned r = 0; for (int l = 0; l < 32; l++) r |= (unsigned)w->i[l] << l; sync(); return r; }
    int scan_incl(int v) { w->i[lane] = v; sync(); int r = 0; for (int l = 0; l <= lane; l++) r += w->i[l]; sync(); return r; }
};

template <int NA, int NF, bool FH, int K>
void run(const GgModel* M, const GgStackTask* T, int n, float* q, float* qd, float* qacc, const float* target, int iterations, float tol,
         int32_t* info) {
    constexpr int NU = NA + NF, NQ = NU + 7 * K, NV = NU + 6 * K;
    auto* In = new gg::StackIn<NA, NF, K>;
    auto* B = new gg::StackBuf<NU, K>;
    HostWarp hw;
    gg::GgChain C; gg::make_chain(*M, C);
    for (int e = 0; e < n; e++) {
        float* pq = q + (size_t)e * NQ; float* pv = qd + (size_t)e * NV; float* pa = qacc + (size_t)e * NV;
        gg::ObjState<K> O;
        for (int k = 0; k < K; k++) {
            const float* o = pq + NU + 7 * k; const float* v = pv + NU + 6 * k;
            O.p[k] = gg::mk(o[0], o[1], o[2]); O.q[k] = gg::mkq(o[3], o[4], o[5], o[6]);
            O.v[k] = gg::mk(v[0], v[1], v[2]); O.w[k] = gg::mk(v[3], v[4], v[5]);
            float* d = In->obj[k];
            for (int i = 0; i < 7; i++) d[i] = o[i];
            for (int i = 0; i < 6; i++) { d[7 + i] = v[i]; In->warm_obj[k][i] = pa[NU + 6 * k + i]; }
        }
        gg::stack_prep<NA, NF, FH, K>(*M, C, pq, pv, pa, target + (size_t)e * NU, O, iterations, tol, true, *In);
        float a[NV];
        for (int i = 0; i < NU; i++) a[i] = In->W.arob[i];
        for (int k = 0; k < K; k++)
            for (int i = 0; i < 6; i++) a[NU + 6 * k + i] = In->aobj[k][i];
        int its[32], nc[32];
        its[0] = 0; nc[0] = (int)In->W.iso_ncon;
        if (In->W.amask != 0.f) {
            std::vector<std::thread> th;
            for (int l = 0; l < 32; l++)
                th.emplace_back([&, l]() {
                    WarpHost wc{l, &hw};
                    gg::stack_solve<NA, NF, FH, K>(*M, T->max_contacts, *In, *B, iterations, tol, wc, a, 1, its[l], nc[l]);
                });
            for (auto& t : th) t.join();
        }
        gg::integrate_robot<NU>(*M, pq, pv, pa, a);
        for (int k = 0; k < K; k++) {
            gg::integrate_cube(*M, O.p[k], O.q[k], O.v[k], O.w[k], pa + NU + 6 * k, a + NU + 6 * k);
            float* o = pq + NU + 7 * k; float* v = pv + NU + 6 * k;
            o[0] = O.p[k].x; o[1] = O.p[k].y; o[2] = O.p[k].z; o[3] = O.q[k].w; o[4] = O.q[k].x; o[5] = O.q[k].y; o[6] = O.q[k].z;
            v[0] = O.v[k].x; v[1] = O.v[k].y; v[2] = O.v[k].z; v[3] = O.w[k].x; v[4] = O.w[k].y; v[5] = O.w[k].z;
        }
        if (info) {
            const int ri = (int)In->W.iso_iters;
            info[e * 4] = nc[0]; info[e * 4 + 1] = its[0] > ri ? its[0] : ri; info[e * 4 + 2] = ((int)In->W.amask) & 1; info[e * 4 + 3] = (int)In->W.amask;
        }
    }
    delete In;
    delete B;
}

template <int NA, int NF, int K>
void obs(const GgModel* M, const GgStackTask* T, int n, const float* q, float* agent, float* envs, float* reward, uint8_t* term) {
    constexpr int NU = NA + NF, NQ = NU + 7 * K;
    for (int e = 0; e < n; e++) {
        const float* pq = q + (size_t)e * NQ;
        gg::ObjState<K> O;
        for (int k = 0; k < K; k++) {
            const float* o = pq + NU + 7 * k;
            O.p[k] = gg::mk(o[0], o[1], o[2]); O.q[k] = gg::mkq(o[3], o[4], o[5], o[6]);
        }
        bool ok;
        gg::stack_observe<NA, NF, K>(*M, *T, pq, O, agent + (size_t)e * (7 + NF), envs + (size_t)e * GG_STACK_ENV_STATE_DIM, reward[e], ok);
        term[e] = ok;
    }
}

}  // namespace

extern "C" {
int hs_sizeof_stack_task(void) { return (int)sizeof(GgStackTask); }
int hs_stack_step(const GgModel* M, const GgStackTask* T, int n, float* q, float* qd, float* qacc, const float* target, int iterations,
                  float tol, int32_t* info) {
    if (M->n_arm == 7 && M->n_fing == 2 && T->n_obj == 5) run<7, 2, false, 5>(M, T, n, q, qd, qacc, target, iterations, tol, info);
    else if (M->n_arm == 5 && M->n_fing == 1 && T->n_obj == 5) run<5, 1, true, 5>(M, T, n, q, qd, qacc, target, iterations, tol, info);
    else return -1;
    return 0;
}
int hs_stack_observe(const GgModel* M, const GgStackTask* T, int n, const float* q, float* agent, float* envs, float* reward, uint8_t* term) {
    if (M->n_arm == 7 && M->n_fing == 2 && T->n_obj == 5) obs<7, 2, 5>(M, T, n, q, agent, envs, reward, term);
    else if (M->n_arm == 5 && M->n_fing == 1 && T->n_obj == 5) obs<5, 1, 5>(M, T, n, q, agent, envs, reward, term);
    else return -1;
    return 0;
}
void hs_stack_spawn(const GgStackTask* T, uint64_t seed, uint32_t gid, uint32_t episode, float* xy) {
    gg::stack_spawn_all(*T, seed, gid, episode, T->n_obj, reinterpret_cast<float(*)[2]>(xy));
}
}

// diagnostic: the cube-cube broad phase (bounding spheres, then the six face axes of the SAT) for n pairs of cube poses (7 floats each)
extern "C" void hs_cubes_may_touch(const GgModel* M, int n, const float* a, const float* b, uint8_t* out) {
    const gg::V3 ch = gg::ld3(M->cube_half);
    for (int i = 0; i < n; i++) {
        const float* p = a + 7 * i; const float* q = b + 7 * i;
        const gg::V3 ca = gg::mk(p[0], p[1], p[2]), cb = gg::mk(q[0], q[1], q[2]);
        const gg::V3 d = ca - cb;
        out[i] = gg::dot(d, d) <= 4.f * gg::dot(ch, ch) &&
                 gg::boxes_may_touch(ca, gg::qmat(gg::mkq(p[3], p[4], p[5], p[6])), ch, cb, gg::qmat(gg::mkq(q[3], q[4], q[5], q[6])), ch);
    }
}
