// =================================================================================================
// ORACLE  --  TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (fp64 and fp32 builds of the same template) of what one `scene.step()` of the
// reference's physics dependency computes for the CubePick scenes, plus the task-level obs /
// reward / reset logic of the reference task classes.
//
//   reference call sites restated here
//     gym_genesis/tasks/franka/cube_pick.py:86-112   reset  (spawn, home pose, hold-home target, 1 settle step)
//     gym_genesis/tasks/franka/cube_pick.py:122-128  step   (control_dofs_position x2, scene.step, reward, obs)
//     gym_genesis/tasks/franka/cube_pick.py:130-135  reward (cube z > 0.1)
//     gym_genesis/tasks/franka/cube_pick.py:137-157  obs    (eef pos/quat, finger q | cube pos/quat, diff, dist)
//     gym_genesis/env.py:61-69                       is_success / terminated / truncated
//
// PARITY UNPINNED.  The arithmetic of `scene.step()` lives in the third-party package
// `genesis-world` (requirement `genesis-world @ git+https://github.com/Genesis-Embodied-AI/Genesis.git`,
// NO tag / commit pin, gym-genesis pyproject.toml:11).  It is not under /root/reference, not
// installed, and not installable (no network); the reference holds no test, fixture or golden
// vector for this path.  This file therefore restates the *published* algorithm family that
// Genesis' rigid solver follows (MuJoCo-lineage soft-constraint dynamics):
//     PD position control with force clamp            (A.3)
//     forward kinematics, joint-space inertia M       (projected Newton-Euler, dense Jacobians)
//     bias forces c(q, qd) incl. gravity              (velocity-product accelerations)
//     Mt = M + armature + dt*(damping + kv)           (approximate implicit-fast, A.3)
//     plane-box (<=4 corners) and box-box (SAT + face clipping, <=8) narrow phase  (A.3)
//     pyramidal friction rows + joint-limit rows with solref/solimp impedance      (A.3/A.4)
//     Newton on the convex cost with exact line search, run to convergence          (A.4)
//     semi-implicit Euler, quaternion integrated on the right by exp(w dt)          (A.3)
// What *is* pinned: the kinematic tree against the published Franka ready-pose TCP and the SO-101
// FK vectors (tests/golden), Philox4x32-10 against the Random123 known-answer vectors, and physics
// identities (M SPD and equal to an independent numpy build, energy / momentum, closed-form free fall
// and resting contact) -- see tests/test_oracle_*.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.
// =================================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>

extern "C" {
typedef struct {
    int32_t nb, nv, nq, nu, ngeom, eef_body, cube_body, max_contacts, n_obj, task;
    const int32_t *parent, *jtype, *dofadr, *qadr, *limited, *ctrl, *geom_body, *geom_group, *obj_body;
    const double *pos, *quat, *axis, *mass, *com, *inertia, *body_invweight, *dof_invweight;
    const double *armature, *damping, *kp, *kv, *frc_lo, *frc_hi, *lim_lo, *lim_hi;
    const double *geom_half, *geom_pos, *geom_quat, *opt, *spawn_quat, *home_q, *obj_spawn, *task_opt;
} OrcModel;
}

namespace {

enum { FIXED = 0, HINGE = 1, SLIDE = 2, FREE = 3 };
enum { OPT_DT = 0, OPT_GRAV, OPT_PLANE_Z, OPT_MU, OPT_TC, OPT_DR, OPT_DMIN, OPT_DMAX, OPT_WIDTH, OPT_MID,
       OPT_POWER, OPT_MINVAL, OPT_SUCCESS_Z, OPT_SPAWN_XLO, OPT_SPAWN_XHI, OPT_SPAWN_YLO, OPT_SPAWN_YHI,
       OPT_SPAWN_Z, OPT_REG_SCALE, OPT_COUNT };

// ------------------------------------------------------------------------------------ vec helpers
template <class T> inline void cross(const T* a, const T* b, T* o) {
    T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
template <class T> inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline void matvec(const T* R, const T* v, T* o) {   // row-major 3x3
    T x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    T y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    T z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
template <class T> inline void matTvec(const T* R, const T* v, T* o) {
    T x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    T y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    T z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
template <class T> inline void matmul3(const T* A, const T* B, T* C) {
    T o[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    for (int i = 0; i < 9; i++) C[i] = o[i];
}
template <class T> inline void quat_mul(const T* a, const T* b, T* o) {   // wxyz
    T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
template <class T> inline void quat_to_mat(const T* q, T* R) {
    T w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}

// ------------------------------------------------------------------------------------ model
template <class T> struct Model {
    int nb, nv, nq, nu, ngeom, eef_body, cube_body, max_contacts;
    int n_obj, task;                      // task 0: pick (franka/cube_pick.py), 1: stack (franka/cube_stack.py)
    std::vector<int> parent, jtype, dofadr, qadr, limited, ctrl, geom_body, geom_group, obj_body;
    std::vector<T> pos, quat, axis, mass, com, inertia, body_w, dof_w, armature, damping, kp, kv, frc_lo, frc_hi,
        lim_lo, lim_hi, geom_half, geom_pos, geom_quat, opt, spawn_quat, home_q;
    std::vector<T> obj_spawn;             // per object: x_lo, x_hi, y_lo, y_hi
    std::vector<T> task_opt;              // stack: xy threshold, z threshold (franka/cube_stack.py:195), min spawn distance
                                          // (so101/cube_stack_batch.py:70), agent_pos = joint positions flag (:176)
    std::vector<int> dof_body;            // body owning each dof
    std::vector<char> affects;            // affects[body*nv + dof]
};

template <class T> static std::vector<T> cvt(const double* p, int n) {
    std::vector<T> v(n);
    for (int i = 0; i < n; i++) v[i] = (T)p[i];
    return v;
}
static std::vector<int> cvti(const int32_t* p, int n) { return std::vector<int>(p, p + n); }

template <class T> static void build(Model<T>& m, const OrcModel* c) {
    m.nb = c->nb; m.nv = c->nv; m.nq = c->nq; m.nu = c->nu; m.ngeom = c->ngeom;
    m.eef_body = c->eef_body; m.cube_body = c->cube_body; m.max_contacts = c->max_contacts;
    m.n_obj = c->n_obj; m.task = c->task;
    m.obj_body = cvti(c->obj_body, m.n_obj); m.obj_spawn = cvt<T>(c->obj_spawn, m.n_obj * 4); m.task_opt = cvt<T>(c->task_opt, 4);
    m.parent = cvti(c->parent, m.nb); m.jtype = cvti(c->jtype, m.nb);
    m.dofadr = cvti(c->dofadr, m.nb); m.qadr = cvti(c->qadr, m.nb);
    m.limited = cvti(c->limited, m.nv); m.ctrl = cvti(c->ctrl, m.nv);
    m.geom_body = cvti(c->geom_body, m.ngeom); m.geom_group = cvti(c->geom_group, m.ngeom);
    m.pos = cvt<T>(c->pos, m.nb * 3); m.quat = cvt<T>(c->quat, m.nb * 4); m.axis = cvt<T>(c->axis, m.nb * 3);
    m.mass = cvt<T>(c->mass, m.nb); m.com = cvt<T>(c->com, m.nb * 3); m.inertia = cvt<T>(c->inertia, m.nb * 6);
    m.body_w = cvt<T>(c->body_invweight, m.nb); m.dof_w = cvt<T>(c->dof_invweight, m.nv);
    m.armature = cvt<T>(c->armature, m.nv); m.damping = cvt<T>(c->damping, m.nv);
    m.kp = cvt<T>(c->kp, m.nv); m.kv = cvt<T>(c->kv, m.nv);
    m.frc_lo = cvt<T>(c->frc_lo, m.nv); m.frc_hi = cvt<T>(c->frc_hi, m.nv);
    m.lim_lo = cvt<T>(c->lim_lo, m.nv); m.lim_hi = cvt<T>(c->lim_hi, m.nv);
    m.geom_half = cvt<T>(c->geom_half, m.ngeom * 3); m.geom_pos = cvt<T>(c->geom_pos, m.ngeom * 3);
    m.geom_quat = cvt<T>(c->geom_quat, m.ngeom * 4);
    m.opt = cvt<T>(c->opt, OPT_COUNT); m.spawn_quat = cvt<T>(c->spawn_quat, 4); m.home_q = cvt<T>(c->home_q, m.nu);
    m.dof_body.assign(m.nv, -1);
    for (int i = 0; i < m.nb; i++) {
        int nd = m.jtype[i] == FREE ? 6 : (m.jtype[i] == FIXED ? 0 : 1);
        for (int k = 0; k < nd; k++) m.dof_body[m.dofadr[i] + k] = i;
    }
    m.affects.assign(m.nb * m.nv, 0);
    for (int i = 0; i < m.nb; i++)
        for (int j = i; j >= 0; j = m.parent[j]) {
            int nd = m.jtype[j] == FREE ? 6 : (m.jtype[j] == FIXED ? 0 : 1);
            for (int k = 0; k < nd; k++) m.affects[i * m.nv + m.dofadr[j] + k] = 1;
        }
}

struct Handle {
    Model<double> d;
    Model<float> f;
};
template <class T> struct Pick;
template <> struct Pick<double> { static const Model<double>& get(const Handle* h) { return h->d; } };
template <> struct Pick<float> { static const Model<float>& get(const Handle* h) { return h->f; } };

// ------------------------------------------------------------------------------------ kinematics
template <class T> struct Kin {
    std::vector<T> xpos, xquat, xmat;          // per body
    std::vector<T> jw, janchor, jv0;           // per dof: angular axis, anchor, constant linear part
    std::vector<T> cpos, Iw;                   // per body: world COM, world inertia (3x3)
    void resize(int nb, int nv) {
        xpos.assign(nb * 3, 0); xquat.assign(nb * 4, 0); xmat.assign(nb * 9, 0);
        jw.assign(nv * 3, 0); janchor.assign(nv * 3, 0); jv0.assign(nv * 3, 0);
        cpos.assign(nb * 3, 0); Iw.assign(nb * 9, 0);
    }
};

// SURVEY.md A.3 "Spatial algebra": free-joint angular dofs are body-frame axes.
static const bool kFreeAngBodyFrame = true;

template <class T> static void forward_kinematics(const Model<T>& m, const T* q, Kin<T>& k) {
    k.resize(m.nb, m.nv);
    for (int i = 0; i < m.nb; i++) {
        T* xp = &k.xpos[i * 3]; T* xq = &k.xquat[i * 4]; T* xm = &k.xmat[i * 9];
        if (m.jtype[i] == FREE) {
            const T* qq = q + m.qadr[i];
            for (int c = 0; c < 3; c++) xp[c] = qq[c];
            for (int c = 0; c < 4; c++) xq[c] = qq[3 + c];
            quat_to_mat(xq, xm);
            int d = m.dofadr[i];
            for (int c = 0; c < 3; c++) {
                k.jv0[(d + c) * 3 + c] = 1;
                for (int r = 0; r < 3; r++) {
                    k.jw[(d + 3 + c) * 3 + r] = kFreeAngBodyFrame ? xm[r * 3 + c] : (T)(r == c);
                    k.janchor[(d + 3 + c) * 3 + r] = xp[r];
                }
            }
        } else {
            T pq[4] = {1, 0, 0, 0}, pp[3] = {0, 0, 0}, pm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            int p = m.parent[i];
            if (p >= 0) {
                for (int c = 0; c < 4; c++) pq[c] = k.xquat[p * 4 + c];
                for (int c = 0; c < 3; c++) pp[c] = k.xpos[p * 3 + c];
                for (int c = 0; c < 9; c++) pm[c] = k.xmat[p * 9 + c];
            }
            T off[3];
            matvec(pm, &m.pos[i * 3], off);
            for (int c = 0; c < 3; c++) xp[c] = pp[c] + off[c];
            quat_mul(pq, &m.quat[i * 4], xq);
            if (m.jtype[i] == HINGE) {
                T ang = q[m.qadr[i]];
                const T* a = &m.axis[i * 3];
                T s = std::sin(ang / 2), c = std::cos(ang / 2);
                T jq[4] = {c, a[0] * s, a[1] * s, a[2] * s}, t[4];
                quat_mul(xq, jq, t);
                for (int e = 0; e < 4; e++) xq[e] = t[e];
            }
            quat_to_mat(xq, xm);
            if (m.jtype[i] == HINGE) {
                int d = m.dofadr[i];
                matvec(xm, &m.axis[i * 3], &k.jw[d * 3]);
                for (int c = 0; c < 3; c++) k.janchor[d * 3 + c] = xp[c];
            } else if (m.jtype[i] == SLIDE) {
                int d = m.dofadr[i];
                T ax[3];
                matvec(xm, &m.axis[i * 3], ax);
                for (int c = 0; c < 3; c++) { k.jv0[d * 3 + c] = ax[c]; xp[c] += ax[c] * q[m.qadr[i]]; }
            }
        }
        T cw[3];
        matvec(xm, &m.com[i * 3], cw);
        for (int c = 0; c < 3; c++) k.cpos[i * 3 + c] = xp[c] + cw[c];
        const T* I6 = &m.inertia[i * 6];
        T Ib[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]}, tmp[9], xmT[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) xmT[r * 3 + c] = xm[c * 3 + r];
        matmul3(xm, Ib, tmp);
        matmul3(tmp, xmT, &k.Iw[i * 9]);
    }
}

// velocity Jacobian column of world point p attached to `body` for dof d (zero if d does not move body)
template <class T> static inline void jac_col(const Model<T>& m, const Kin<T>& k, int body, int d, const T* p, T* jv, T* jw) {
    if (body < 0 || !m.affects[body * m.nv + d]) { jv[0] = jv[1] = jv[2] = 0; if (jw) jw[0] = jw[1] = jw[2] = 0; return; }
    T r[3] = {p[0] - k.janchor[d * 3], p[1] - k.janchor[d * 3 + 1], p[2] - k.janchor[d * 3 + 2]};
    cross(&k.jw[d * 3], r, jv);
    for (int c = 0; c < 3; c++) jv[c] += k.jv0[d * 3 + c];
    if (jw) for (int c = 0; c < 3; c++) jw[c] = k.jw[d * 3 + c];
}

// ------------------------------------------------------------------------------------ dynamics
template <class T> struct Dyn {
    std::vector<T> M, Mt, bias, tau, fs, as;
};

template <class T> static void dynamics(const Model<T>& m, const Kin<T>& k, const T* q, const T* qd, const T* target, Dyn<T>& d) {
    const int nv = m.nv, nb = m.nb;
    d.M.assign(nv * nv, 0); d.Mt.assign(nv * nv, 0); d.bias.assign(nv, 0); d.tau.assign(nv, 0); d.fs.assign(nv, 0); d.as.assign(nv, 0);
    // --- joint-space inertia: M = sum_b m Jv^T Jv + Jw^T Iw Jw
    std::vector<T> Jv(3 * nv), Jw(3 * nv), IJw(3 * nv);
    // --- body velocities / velocity-product accelerations (qdd = 0)
    std::vector<T> w(nb * 3, 0), al(nb * 3, 0), vo(nb * 3, 0), ao(nb * 3, 0);
    const T g = m.opt[OPT_GRAV];
    for (int i = 0; i < nb; i++) {
        T* wi = &w[i * 3]; T* ali = &al[i * 3]; T* voi = &vo[i * 3]; T* aoi = &ao[i * 3];
        if (m.jtype[i] == FREE) {
            int dd = m.dofadr[i];
            for (int c = 0; c < 3; c++) voi[c] = qd[dd + c];
            T wb[3] = {qd[dd + 3], qd[dd + 4], qd[dd + 5]};
            if (kFreeAngBodyFrame) matvec(&k.xmat[i * 9], wb, wi); else for (int c = 0; c < 3; c++) wi[c] = wb[c];
            // body-frame axes rotate with the body: d/dt(R wb) = w x (R wb) = 0  -> no bias acceleration
        } else {
            int p = m.parent[i];
            T wp[3] = {0, 0, 0}, alp[3] = {0, 0, 0}, vop[3] = {0, 0, 0}, aop[3] = {0, 0, 0}, op[3] = {0, 0, 0};
            if (p >= 0) for (int c = 0; c < 3; c++) { wp[c] = w[p * 3 + c]; alp[c] = al[p * 3 + c]; vop[c] = vo[p * 3 + c]; aop[c] = ao[p * 3 + c]; op[c] = k.xpos[p * 3 + c]; }
            T r[3] = {k.xpos[i * 3] - op[0], k.xpos[i * 3 + 1] - op[1], k.xpos[i * 3 + 2] - op[2]};
            T wr[3], wwr[3], ar[3];
            cross(wp, r, wr); cross(wp, wr, wwr); cross(alp, r, ar);
            for (int c = 0; c < 3; c++) { wi[c] = wp[c]; ali[c] = alp[c]; voi[c] = vop[c] + wr[c]; aoi[c] = aop[c] + ar[c] + wwr[c]; }
            if (m.jtype[i] == HINGE) {
                int dd = m.dofadr[i];
                const T* kx = &k.jw[dd * 3];
                T wk[3];
                cross(wp, kx, wk);
                for (int c = 0; c < 3; c++) { wi[c] += kx[c] * qd[dd]; ali[c] += wk[c] * qd[dd]; }
            } else if (m.jtype[i] == SLIDE) {
                int dd = m.dofadr[i];
                const T* kx = &k.jv0[dd * 3];
                T wk[3];
                cross(wp, kx, wk);
                for (int c = 0; c < 3; c++) { voi[c] += kx[c] * qd[dd]; aoi[c] += 2 * wk[c] * qd[dd]; }
            }
        }
    }
    for (int i = 0; i < nb; i++) {
        if (m.mass[i] <= 0) continue;
        const T* c = &k.cpos[i * 3];
        for (int dd = 0; dd < nv; dd++) {
            T jv[3], jw[3];
            jac_col(m, k, i, dd, c, jv, jw);
            for (int r = 0; r < 3; r++) { Jv[r * nv + dd] = jv[r]; Jw[r * nv + dd] = jw[r]; }
        }
        const T* I = &k.Iw[i * 9];
        for (int dd = 0; dd < nv; dd++)
            for (int r = 0; r < 3; r++) IJw[r * nv + dd] = I[r * 3] * Jw[dd] + I[r * 3 + 1] * Jw[nv + dd] + I[r * 3 + 2] * Jw[2 * nv + dd];
        for (int a = 0; a < nv; a++) {
            if (!m.affects[i * nv + a]) continue;
            for (int b = 0; b < nv; b++) {
                if (!m.affects[i * nv + b]) continue;
                T s = 0;
                for (int r = 0; r < 3; r++) s += m.mass[i] * Jv[r * nv + a] * Jv[r * nv + b] + Jw[r * nv + a] * IJw[r * nv + b];
                d.M[a * nv + b] += s;
            }
        }
        // bias: COM acceleration (velocity products) minus gravity, and Euler's equation
        T rc[3] = {c[0] - k.xpos[i * 3], c[1] - k.xpos[i * 3 + 1], c[2] - k.xpos[i * 3 + 2]};
        const T* wi = &w[i * 3]; const T* ali = &al[i * 3];
        T wr[3], wwr[3], ar[3], Iwv[3], wIw[3], Ial[3];
        cross(wi, rc, wr); cross(wi, wr, wwr); cross(ali, rc, ar);
        T f[3] = {m.mass[i] * (ao[i * 3] + ar[0] + wwr[0]), m.mass[i] * (ao[i * 3 + 1] + ar[1] + wwr[1]),
                  m.mass[i] * (ao[i * 3 + 2] + ar[2] + wwr[2] + g)};
        matvec(I, wi, Iwv); cross(wi, Iwv, wIw); matvec(I, ali, Ial);
        T n[3] = {Ial[0] + wIw[0], Ial[1] + wIw[1], Ial[2] + wIw[2]};
        for (int dd = 0; dd < nv; dd++) {
            if (!m.affects[i * nv + dd]) continue;
            T s = 0;
            for (int r = 0; r < 3; r++) s += Jv[r * nv + dd] * f[r] + Jw[r * nv + dd] * n[r];
            d.bias[dd] += s;
        }
    }
    // --- PD position control (A.3): tau = clamp(kp (target - q) - kv qd) - damping qd
    const T dt = m.opt[OPT_DT];
    for (int dd = 0; dd < nv; dd++) {
        T t = 0;
        if (m.ctrl[dd]) {
            int b = m.dof_body[dd];
            t = m.kp[dd] * (target[dd] - q[m.qadr[b]]) - m.kv[dd] * qd[dd];
            t = std::min(std::max(t, m.frc_lo[dd]), m.frc_hi[dd]);
        }
        d.tau[dd] = t - m.damping[dd] * qd[dd];
        d.fs[dd] = d.tau[dd] - d.bias[dd];
    }
    for (int a = 0; a < nv; a++) {
        d.M[a * nv + a] += m.armature[a];
        for (int b = 0; b < nv; b++) d.Mt[a * nv + b] = d.M[a * nv + b];
        d.Mt[a * nv + a] += dt * (m.damping[a] + (m.ctrl[a] ? m.kv[a] : (T)0));
    }
}

// dense Cholesky (lower), in place; returns false if not PD
template <class T> static bool chol(std::vector<T>& A, int n) {
    for (int j = 0; j < n; j++) {
        T s = A[j * n + j];
        for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
        if (!(s > 0)) return false;
        T l = std::sqrt(s);
        A[j * n + j] = l;
        for (int i = j + 1; i < n; i++) {
            T t = A[i * n + j];
            for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / l;
        }
    }
    return true;
}
template <class T> static void chol_solve(const std::vector<T>& L, int n, const T* b, T* x) {
    std::vector<T> y(n);
    for (int i = 0; i < n; i++) { T s = b[i]; for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k]; y[i] = s / L[i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { T s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k]; x[i] = s / L[i * n + i]; }
}

// ------------------------------------------------------------------------------------ collision
template <class T> struct Contact {
    T p[3], n[3], depth;      // n points from body a to body b
    int ba, bb;               // -1 = world
};

// plane z = h (normal +z) vs box: first <=4 corners with signed distance <= 0 (A.3, "mjc_PlaneBox")
template <class T> static void plane_box(T h, const T* c, const T* R, const T* half, int body, int maxc, std::vector<Contact<T>>& out) {
    int cnt = 0;
    for (int i = 0; i < 8 && cnt < 4 && (int)out.size() < maxc; i++) {
        T l[3] = {(i & 1) ? half[0] : -half[0], (i & 2) ? half[1] : -half[1], (i & 4) ? half[2] : -half[2]};
        T w[3];
        matvec(R, l, w);
        T corner[3] = {c[0] + w[0], c[1] + w[1], c[2] + w[2]};
        T dist = corner[2] - h;
        if (dist > 0) continue;
        Contact<T> ct;
        ct.p[0] = corner[0]; ct.p[1] = corner[1]; ct.p[2] = corner[2] - dist / 2;
        ct.n[0] = 0; ct.n[1] = 0; ct.n[2] = 1;
        ct.depth = -dist; ct.ba = -1; ct.bb = body;
        out.push_back(ct);
        cnt++;
    }
}

template <class T> static inline T sgn1(T x) { return x < 0 ? (T)-1 : (T)1; }

// clip polygon (u,v,w coords) against  s*coord[ax] <= lim
template <class T> static int clip_poly(const T (*in)[3], int n, T (*out)[3], int ax, T s, T lim) {
    int m = 0;
    for (int i = 0; i < n; i++) {
        const T* a = in[i]; const T* b = in[(i + 1) % n];
        T da = s * a[ax] - lim, db = s * b[ax] - lim;
        if (da <= 0) { if (m < 8) { out[m][0] = a[0]; out[m][1] = a[1]; out[m][2] = a[2]; m++; } }
        if ((da <= 0) != (db <= 0)) {
            T t = da / (da - db);
            if (m < 8) { for (int c = 0; c < 3; c++) out[m][c] = a[c] + t * (b[c] - a[c]); m++; }
        }
    }
    return m;
}

// box A (body ba) vs box B (body bb): SAT over 15 axes, face clipping / edge-edge closest points.
// Axis preference: A faces (strict <), then B faces and edge axes must win by a 5 % margin.
template <class T> static void box_box(const T* cA, const T* RA, const T* hA, int ba, const T* cB, const T* RB, const T* hB, int bb,
                                        int maxc, std::vector<Contact<T>>& out) {
    const T FUDGE = (T)1.05;
    T A[3][3], B[3][3];           // A[i] = i-th axis of box A in world
    for (int i = 0; i < 3; i++) for (int r = 0; r < 3; r++) { A[i][r] = RA[r * 3 + i]; B[i][r] = RB[r * 3 + i]; }
    T t[3] = {cB[0] - cA[0], cB[1] - cA[1], cB[2] - cA[2]};
    T C[3][3], aC[3][3], tA[3], tB[3];
    for (int i = 0; i < 3; i++) { tA[i] = dot3(t, A[i]); tB[i] = dot3(t, B[i]); for (int j = 0; j < 3; j++) { C[i][j] = dot3(A[i], B[j]); aC[i][j] = std::fabs(C[i][j]); } }
    int best = -1; T bestpen = 0;
    for (int i = 0; i < 3; i++) {
        T pen = hA[i] + hB[0] * aC[i][0] + hB[1] * aC[i][1] + hB[2] * aC[i][2] - std::fabs(tA[i]);
        if (pen < 0) return;
        if (best < 0 || pen < bestpen) { best = i; bestpen = pen; }
    }
    for (int j = 0; j < 3; j++) {
        T pen = hB[j] + hA[0] * aC[0][j] + hA[1] * aC[1][j] + hA[2] * aC[2][j] - std::fabs(tB[j]);
        if (pen < 0) return;
        if (pen * FUDGE < bestpen) { best = 3 + j; bestpen = pen; }
    }
    T bestL[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            T L[3];
            cross(A[i], B[j], L);
            T l2 = dot3(L, L);
            if (l2 < (T)1e-6) continue;
            T l = std::sqrt(l2);
            int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            T rA = hA[i1] * aC[i2][j] + hA[i2] * aC[i1][j];
            T rB = hB[j1] * aC[i][j2] + hB[j2] * aC[i][j1];
            T pen = (rA + rB - std::fabs(dot3(t, L))) / l;
            if (pen < 0) return;
            if (pen * FUDGE < bestpen) { best = 6 + i * 3 + j; bestpen = pen; for (int c = 0; c < 3; c++) bestL[c] = L[c] / l; }
        }
    if (best >= 6) {
        // ---- edge-edge: one contact at the midpoint of the closest points of the two supporting edges
        int i = (best - 6) / 3, j = (best - 6) % 3;
        T n[3];
        T s = sgn1(dot3(t, bestL));
        for (int c = 0; c < 3; c++) n[c] = s * bestL[c];
        T pA[3] = {cA[0], cA[1], cA[2]}, pB[3] = {cB[0], cB[1], cB[2]};
        for (int k = 0; k < 3; k++) {
            if (k != i) { T sg = sgn1(dot3(n, A[k])); for (int c = 0; c < 3; c++) pA[c] += sg * hA[k] * A[k][c]; }
            if (k != j) { T sg = sgn1(dot3(n, B[k])); for (int c = 0; c < 3; c++) pB[c] -= sg * hB[k] * B[k][c]; }
        }
        T d[3] = {pB[0] - pA[0], pB[1] - pA[1], pB[2] - pA[2]};
        T ab = C[i][j], da = dot3(d, A[i]), db = dot3(d, B[j]);
        T den = 1 - ab * ab;
        T sa = (da - ab * db) / den, sb = (ab * da - db) / den;
        Contact<T> ct;
        for (int c = 0; c < 3; c++) { ct.p[c] = (T)0.5 * (pA[c] + sa * A[i][c] + pB[c] + sb * B[j][c]); ct.n[c] = n[c]; }
        ct.depth = bestpen; ct.ba = ba; ct.bb = bb;
        if ((int)out.size() < maxc) out.push_back(ct);
        return;
    }
    // ---- face contact: reference box = owner of the best face axis, incident = the other box
    const bool refA = best < 3;
    const int ax = refA ? best : best - 3;
    const T (*Rf)[3] = refA ? A : B; const T (*In)[3] = refA ? B : A;
    const T* cR = refA ? cA : cB; const T* cI = refA ? cB : cA;
    const T* hR = refA ? hA : hB; const T* hI = refA ? hB : hA;
    T tr[3] = {cI[0] - cR[0], cI[1] - cR[1], cI[2] - cR[2]};
    T sg = sgn1(dot3(tr, Rf[ax]));
    T nref[3] = {sg * Rf[ax][0], sg * Rf[ax][1], sg * Rf[ax][2]};
    // incident face: most anti-parallel to nref
    int jf = 0; T bd = -1;
    T dj[3];
    for (int j = 0; j < 3; j++) { dj[j] = dot3(nref, In[j]); if (std::fabs(dj[j]) > bd) { bd = std::fabs(dj[j]); jf = j; } }
    T fs = -sgn1(dj[jf]);
    int j1 = (jf + 1) % 3, j2 = (jf + 2) % 3, u = (ax + 1) % 3, v = (ax + 2) % 3;
    static const int S1[4] = {1, -1, -1, 1}, S2[4] = {1, 1, -1, -1};
    T poly[2][8][3];
    for (int k = 0; k < 4; k++) {
        T wv[3];
        for (int c = 0; c < 3; c++) wv[c] = cI[c] + fs * hI[jf] * In[jf][c] + S1[k] * hI[j1] * In[j1][c] + S2[k] * hI[j2] * In[j2][c] - cR[c];
        poly[0][k][0] = dot3(wv, Rf[u]); poly[0][k][1] = dot3(wv, Rf[v]); poly[0][k][2] = dot3(wv, nref);
    }
    int n = 4;
    n = clip_poly(poly[0], n, poly[1], 0, (T)1, hR[u]);
    n = clip_poly(poly[1], n, poly[0], 0, (T)-1, hR[u]);
    n = clip_poly(poly[0], n, poly[1], 1, (T)1, hR[v]);
    n = clip_poly(poly[1], n, poly[0], 1, (T)-1, hR[v]);
    for (int k = 0; k < n && (int)out.size() < maxc; k++) {
        T depth = hR[ax] - poly[0][k][2];
        if (depth < 0) continue;
        Contact<T> ct;
        T wn = poly[0][k][2] + depth / 2;
        for (int c = 0; c < 3; c++) {
            ct.p[c] = cR[c] + poly[0][k][0] * Rf[u][c] + poly[0][k][1] * Rf[v][c] + wn * nref[c];
            ct.n[c] = refA ? nref[c] : -nref[c];
        }
        ct.depth = depth; ct.ba = ba; ct.bb = bb;
        out.push_back(ct);
    }
}

template <class T> static void geom_pose(const Model<T>& m, const Kin<T>& k, int g, T* c, T* R) {
    int b = m.geom_body[g];
    T Rg[9], off[3];
    quat_to_mat(&m.geom_quat[g * 4], Rg);
    matmul3(&k.xmat[b * 9], Rg, R);
    matvec(&k.xmat[b * 9], &m.geom_pos[g * 3], off);
    for (int i = 0; i < 3; i++) c[i] = k.xpos[b * 3 + i] + off[i];
}

// Order (deterministic): objects-plane, pads-plane in geom order, pad-object (pad-major, object-minor), object-object
// pairs (i < j, row-major) -- the cube-cube box-box pairs of franka/cube_stack.py (RigidOptions(box_box_detection=True),
// :46); capped at max_contacts.
template <class T> static void collide(const Model<T>& m, const Kin<T>& k, std::vector<Contact<T>>& out) {
    out.clear();
    const T h = m.opt[OPT_PLANE_Z];
    T c[3], R[9], c2[3], R2[9];
    for (int pass = 0; pass < 2; pass++)
        for (int g = 0; g < m.ngeom; g++) {
            if ((m.geom_group[g] == 1) != (pass == 0)) continue;
            geom_pose(m, k, g, c, R);
            plane_box(h, c, R, &m.geom_half[g * 3], m.geom_body[g], m.max_contacts, out);
        }
    for (int g = 0; g < m.ngeom; g++) {
        if (m.geom_group[g] != 0) continue;
        geom_pose(m, k, g, c, R);
        for (int o = 0; o < m.ngeom; o++) {
            if (m.geom_group[o] != 1) continue;
            geom_pose(m, k, o, c2, R2);
            box_box(c, R, &m.geom_half[g * 3], m.geom_body[g], c2, R2, &m.geom_half[o * 3], m.geom_body[o], m.max_contacts, out);
        }
    }
    for (int g = 0; g < m.ngeom; g++) {
        if (m.geom_group[g] != 1) continue;
        geom_pose(m, k, g, c, R);
        for (int o = g + 1; o < m.ngeom; o++) {
            if (m.geom_group[o] != 1 || m.geom_body[o] == m.geom_body[g]) continue;
            geom_pose(m, k, o, c2, R2);
            box_box(c, R, &m.geom_half[g * 3], m.geom_body[g], c2, R2, &m.geom_half[o * 3], m.geom_body[o], m.max_contacts, out);
        }
    }
}

// deterministic tangent basis of a unit normal
template <class T> static void tangent_basis(const T* n, T* t1, T* t2) {
    T ref[3] = {0, 0, 0};
    if (std::fabs(n[0]) < (T)0.57735) ref[0] = 1; else ref[1] = 1;
    T d = dot3(ref, n);
    for (int c = 0; c < 3; c++) t1[c] = ref[c] - d * n[c];
    T l = std::sqrt(dot3(t1, t1));
    for (int c = 0; c < 3; c++) t1[c] /= l;
    cross(n, t1, t2);
}

// ------------------------------------------------------------------------------------ constraints
template <class T> struct Rows {
    int n = 0;
    std::vector<T> J, aref, D;    // J row-major n x nv
};

template <class T> static void imp_aref(const Model<T>& m, T pos, T vel, T& imp, T& aref) {
    const T tc = m.opt[OPT_TC], dr = m.opt[OPT_DR], dmin = m.opt[OPT_DMIN], dmax = m.opt[OPT_DMAX];
    const T width = m.opt[OPT_WIDTH], mid = m.opt[OPT_MID], power = m.opt[OPT_POWER];
    T x = std::fabs(pos) / width;
    if (x >= 1) imp = dmax;
    else if (x <= 0) imp = dmin;
    else {
        T y = x <= mid ? std::pow(x, power) / std::pow(mid, power - 1) : 1 - std::pow(1 - x, power) / std::pow(1 - mid, power - 1);
        imp = dmin + y * (dmax - dmin);
    }
    T Bc = 2 / (dmax * tc), Kc = 1 / (dmax * dmax * tc * tc * dr * dr);
    aref = -Bc * vel - Kc * imp * pos;
}

template <class T> static void make_rows(const Model<T>& m, const Kin<T>& k, const T* q, const T* qd, const std::vector<Contact<T>>& cons, Rows<T>& r) {
    const int nv = m.nv;
    r.n = 0; r.J.clear(); r.aref.clear(); r.D.clear();
    const T minval = m.opt[OPT_MINVAL];
    // joint limits (A.4): pos = q - lo (J=+1) or hi - q (J=-1), row only if pos < 0
    for (int d = 0; d < nv; d++) {
        if (!m.limited[d]) continue;
        T qq = q[m.qadr[m.dof_body[d]]];
        for (int side = 0; side < 2; side++) {
            T pos = side == 0 ? qq - m.lim_lo[d] : m.lim_hi[d] - qq;
            if (!(pos < 0)) continue;
            T sgn = side == 0 ? 1 : -1;
            T imp, aref;
            imp_aref(m, pos, sgn * qd[d], imp, aref);
            T R = std::max(minval, (1 - imp) / imp * m.dof_w[d]);
            r.J.resize((r.n + 1) * nv, 0);
            r.J[r.n * nv + d] = sgn;
            r.aref.push_back(aref); r.D.push_back(1 / R); r.n++;
        }
    }
    // contacts: 4 pyramidal rows each, dir = n + s*mu*t_k with n from a to b and J = J_b - J_a, so that
    // J qd is the separation rate (A.3/A.4; SURVEY writes the same rows with the opposite normal convention)
    const T mu = m.opt[OPT_MU];
    std::vector<T> Jn(nv), Jt1(nv), Jt2(nv);
    for (const auto& c : cons) {
        T t1[3], t2[3];
        tangent_basis(c.n, t1, t2);
        for (int d = 0; d < nv; d++) {
            T va[3], vb[3];
            jac_col(m, k, c.ba, d, c.p, va, (T*)nullptr);
            jac_col(m, k, c.bb, d, c.p, vb, (T*)nullptr);
            T rel[3] = {vb[0] - va[0], vb[1] - va[1], vb[2] - va[2]};
            Jn[d] = dot3(c.n, rel); Jt1[d] = dot3(t1, rel); Jt2[d] = dot3(t2, rel);
        }
        T wa = c.ba >= 0 ? m.body_w[c.ba] : 0, wb = c.bb >= 0 ? m.body_w[c.bb] : 0;
        for (int kk = 0; kk < 4; kk++) {
            const std::vector<T>& Jt = kk < 2 ? Jt1 : Jt2;
            T s = (kk & 1) ? 1 : -1;
            r.J.resize((r.n + 1) * nv, 0);
            T vel = 0;
            for (int d = 0; d < nv; d++) { T j = Jn[d] + s * mu * Jt[d]; r.J[r.n * nv + d] = j; vel += j * qd[d]; }
            T imp, aref;
            imp_aref(m, -c.depth, vel, imp, aref);
            T R = std::max(minval, (T)m.opt[OPT_REG_SCALE] * (1 - imp) / imp * (wa + wb));   // GENESIS_COMPAT_PYRAMID_MU_SCALING: 2 mu^2 (1 + mu^2) or 1
            r.aref.push_back(aref); r.D.push_back(1 / R); r.n++;
        }
    }
}

// ------------------------------------------------------------------------------------ Newton solver (A.4)
template <class T> struct SolveInfo { int iters; T grad; T cost; };

template <class T> static T eval_cost(int nv, const std::vector<T>& Mt, const T* as, const Rows<T>& r, const T* a) {
    T c = 0;
    std::vector<T> e(nv);
    for (int i = 0; i < nv; i++) e[i] = a[i] - as[i];
    for (int i = 0; i < nv; i++) { T s = 0; for (int j = 0; j < nv; j++) s += Mt[i * nv + j] * e[j]; c += (T)0.5 * e[i] * s; }
    for (int i = 0; i < r.n; i++) { T jar = -r.aref[i]; for (int j = 0; j < nv; j++) jar += r.J[i * nv + j] * a[j]; if (jar < 0) c += (T)0.5 * r.D[i] * jar * jar; }
    return c;
}

template <class T> static SolveInfo<T> newton(int nv, const std::vector<T>& Mt, const T* as, const Rows<T>& r, const T* warm, T* a, int iterations, T tol) {
    SolveInfo<T> info{0, 0, 0};
    // warm start: previous qacc if cheaper than the unconstrained solution (A.4)
    T cw = eval_cost(nv, Mt, as, r, warm), cs = eval_cost(nv, Mt, as, r, as);
    for (int i = 0; i < nv; i++) a[i] = cw < cs ? warm[i] : as[i];
    if (r.n == 0) { for (int i = 0; i < nv; i++) a[i] = as[i]; return info; }
    T meaninertia = 0;
    for (int i = 0; i < nv; i++) meaninertia += Mt[i * nv + i];
    meaninertia /= nv;
    const T scale = 1 / (meaninertia * std::max(1, nv));
    std::vector<T> jar(r.n), jv(r.n), g(nv), s(nv), H(nv * nv), Ms(nv), e(nv);
    for (int it = 0; it < iterations; it++) {
        for (int i = 0; i < r.n; i++) { T v = -r.aref[i]; for (int j = 0; j < nv; j++) v += r.J[i * nv + j] * a[j]; jar[i] = v; }
        for (int i = 0; i < nv; i++) e[i] = a[i] - as[i];
        for (int i = 0; i < nv; i++) { T v = 0; for (int j = 0; j < nv; j++) v += Mt[i * nv + j] * e[j]; g[i] = v; }
        H = Mt;
        for (int i = 0; i < r.n; i++) {
            if (!(jar[i] < 0)) continue;
            const T* Ji = &r.J[i * nv];
            for (int x = 0; x < nv; x++) {
                g[x] += r.D[i] * jar[i] * Ji[x];
                if (Ji[x] == 0) continue;
                for (int y = 0; y < nv; y++) H[x * nv + y] += r.D[i] * Ji[x] * Ji[y];
            }
        }
        T gn = 0;
        for (int i = 0; i < nv; i++) gn += g[i] * g[i];
        gn = std::sqrt(gn);
        info.grad = gn * scale; info.iters = it;
        if (gn * scale < tol) break;
        if (!chol(H, nv)) break;
        std::vector<T> mg(nv);
        for (int i = 0; i < nv; i++) mg[i] = -g[i];
        chol_solve(H, nv, mg.data(), s.data());
        // exact line search on the piecewise-quadratic phi(alpha) = cost(a + alpha s)
        T q1 = 0, q2 = 0;
        for (int i = 0; i < nv; i++) { T v = 0; for (int j = 0; j < nv; j++) v += Mt[i * nv + j] * s[j]; Ms[i] = v; }
        for (int i = 0; i < nv; i++) { q1 += Ms[i] * e[i]; q2 += Ms[i] * s[i]; }
        for (int i = 0; i < r.n; i++) { T v = 0; for (int j = 0; j < nv; j++) v += r.J[i * nv + j] * s[j]; jv[i] = v; }
        auto dphi = [&](T al, T& d2) {
            T d1 = q1 + al * q2; d2 = q2;
            for (int i = 0; i < r.n; i++) { T x = jar[i] + al * jv[i]; if (x < 0) { d1 += r.D[i] * x * jv[i]; d2 += r.D[i] * jv[i] * jv[i]; } }
            return d1;
        };
        T d2, d0 = dphi(0, d2);
        T al = 0, lo = 0, hi = -1, d1 = d0;
        for (int ls = 0; ls < 60; ls++) {
            T nx = al - d1 / d2;
            if (hi >= 0 && (nx <= lo || nx >= hi)) nx = (T)0.5 * (lo + hi);
            if (nx <= 0) nx = hi >= 0 ? (T)0.5 * (lo + hi) : (T)1e-3;
            al = nx;
            d1 = dphi(al, d2);
            if (d1 < 0) lo = al; else hi = al;
            if (std::fabs(d1) <= std::fabs(d0) * (T)1e-12) break;
        }
        for (int i = 0; i < nv; i++) a[i] += al * s[i];
        info.iters = it + 1;
    }
    info.cost = eval_cost(nv, Mt, as, r, a);
    return info;
}

// ------------------------------------------------------------------------------------ one substep
template <class T> struct StepOut { int ncon, nrows, iters; T grad; };

template <class T> static StepOut<T> substep(const Model<T>& m, T* q, T* qd, T* qacc, const T* target_u, int iterations, T tol) {
    const int nv = m.nv;
    Kin<T> k; Dyn<T> d; std::vector<Contact<T>> cons; Rows<T> rows;
    std::vector<T> target(nv, 0);
    for (int i = 0; i < m.nu; i++) target[i] = target_u[i];
    forward_kinematics(m, q, k);
    dynamics(m, k, q, qd, target.data(), d);
    std::vector<T> L = d.Mt;
    chol(L, nv);
    chol_solve(L, nv, d.fs.data(), d.as.data());
    collide(m, k, cons);
    make_rows(m, k, q, qd, cons, rows);
    std::vector<T> a(nv);
    SolveInfo<T> si = newton(nv, d.Mt, d.as.data(), rows, qacc, a.data(), iterations, tol);
    // semi-implicit Euler (A.3)
    const T dt = m.opt[OPT_DT];
    for (int i = 0; i < nv; i++) { qacc[i] = a[i]; qd[i] += a[i] * dt; }
    for (int b = 0; b < m.nb; b++) {
        if (m.jtype[b] == HINGE || m.jtype[b] == SLIDE) q[m.qadr[b]] += qd[m.dofadr[b]] * dt;
        else if (m.jtype[b] == FREE) {
            T* qq = q + m.qadr[b]; const T* v = qd + m.dofadr[b];
            for (int c = 0; c < 3; c++) qq[c] += v[c] * dt;
            T w[3] = {v[3] * dt, v[4] * dt, v[5] * dt};
            T ang = std::sqrt(dot3(w, w));
            T dq[4] = {1, 0, 0, 0};
            if (ang > 0) { T s = std::sin(ang / 2) / ang; dq[0] = std::cos(ang / 2); dq[1] = w[0] * s; dq[2] = w[1] * s; dq[3] = w[2] * s; }
            T nq[4];
            if (kFreeAngBodyFrame) quat_mul(qq + 3, dq, nq); else quat_mul(dq, qq + 3, nq);
            T l = std::sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
            for (int c = 0; c < 4; c++) qq[3 + c] = nq[c] / l;
        }
    }
    StepOut<T> o{(int)cons.size(), rows.n, si.iters, si.grad};
    return o;
}

template <class T> static void observe(const Model<T>& m, const T* q, T* agent, int agent_dim, T* envs, T* reward, uint8_t* term) {
    Kin<T> k;
    forward_kinematics(m, q, k);
    const int e = m.eef_body, cb = m.cube_body, nf = agent_dim - 7;
    const T* ep = &k.xpos[e * 3]; const T* eq = &k.xquat[e * 4];
    const T* cp = &k.xpos[cb * 3]; const T* cq = &k.xquat[cb * 4];
    if (m.task == 1 && m.task_opt[3] != 0) {
        for (int c = 0; c < m.nu; c++) agent[c] = q[c];          // so101/cube_stack_batch.py:176: agent_pos = get_qpos()
    } else {
        for (int c = 0; c < 3; c++) agent[c] = ep[c];
        for (int c = 0; c < 4; c++) agent[3 + c] = eq[c];
        for (int c = 0; c < nf; c++) agent[7 + c] = q[m.nu - nf + c];
    }
    for (int c = 0; c < 3; c++) envs[c] = cp[c];
    for (int c = 0; c < 4; c++) envs[3 + c] = cq[c];
    T dd[3] = {ep[0] - cp[0], ep[1] - cp[1], ep[2] - cp[2]};
    for (int c = 0; c < 3; c++) envs[7 + c] = dd[c];
    envs[10] = std::sqrt(dot3(dd, dd));
    // reference compares a float32 z with 0.1 under NumPy weak-scalar promotion => z > 0.1f
    bool succ = (float)cp[2] > (float)m.opt[OPT_SUCCESS_Z];
    if (m.task == 1) {
        // franka/cube_stack.py:188-216: environment_state = [cube1 pos, cube1 quat, eef - cube1, |.|, cube2 pos] (14);
        // reward = (|xy1 - xy2| < 0.05) & (z1 - z2 > 0.03) evaluated on float32 device tensors (torch weak-scalar promotion
        // keeps float32).  The 2-norm is restated as sqrtf(fmaf(dx, dx, dy * dy)); the CUDA path uses the same expression.
        const T* c2 = &k.xpos[m.obj_body[1] * 3];
        for (int c = 0; c < 3; c++) envs[11 + c] = c2[c];
        const float dx = (float)cp[0] - (float)c2[0], dy = (float)cp[1] - (float)c2[1];
        const float xy = std::sqrt(std::fmaf(dx, dx, dy * dy));
        const float dz = (float)cp[2] - (float)c2[2];
        succ = xy < (float)m.task_opt[0] && dz > (float)m.task_opt[1];
    }
    *reward = succ ? 1 : 0;
    *term = succ ? 1 : 0;
}

// ------------------------------------------------------------------------------------ inverse kinematics
// Restates robot.inverse_kinematics(link=eef, pos=, quat=) as the reference's expert policies call it
// (examples/franka/pick_cube_state.py:46-51): Genesis' damped-least-squares iteration [UPSTREAM-RECALL]
//     e = [target_pos - pos ; rotvec(target_quat * conj(quat))] (masked),  solved: |e_pos| < pos_tol and |e_rot| < rot_tol
//     dq = J^T (J J^T + damping^2 I)^-1 e over the dofs that move the link, clamp(|dq_i| <= max_step), clip to joint limits
// first sample only (init = given q).  Dense Jacobian columns + dense 6x6 Cholesky: deliberately NOT the shifted
// Gram-matrix formulation the CUDA kernel uses (gym_genesis_b200/csrc/gg_ik.cuh).
template <class T> static int inverse_kinematics(const Model<T>& m, T* qfull, const T* tpos, const T* tquat, int max_iters, T damping,
                                                 T pos_tol, T rot_tol, T max_step, int respect_limits, const T* mask6, T* err2) {
    Kin<T> k;
    int used = 0;
    const int b = m.eef_body;
    for (int it = 0;; it++) {
        forward_kinematics(m, qfull, k);
        const T* xp = &k.xpos[b * 3]; const T* xq = &k.xquat[b * 4];
        T e[6] = {0, 0, 0, 0, 0, 0};
        if (tpos) for (int c = 0; c < 3; c++) e[c] = (tpos[c] - xp[c]) * mask6[c];
        if (tquat) {
            T inv[4] = {xq[0], -xq[1], -xq[2], -xq[3]}, qe[4];
            quat_mul(tquat, inv, qe);
            T sg = qe[0] < 0 ? (T)-1 : (T)1;
            T s = std::sqrt(qe[1] * qe[1] + qe[2] * qe[2] + qe[3] * qe[3]);
            T kk = 2 * sg;
            if (s > (T)1e-9) kk = sg * 2 * std::atan2(s, std::fabs(qe[0])) / s;
            for (int c = 0; c < 3; c++) e[3 + c] = qe[1 + c] * kk * mask6[3 + c];
        }
        err2[0] = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        err2[1] = std::sqrt(e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
        if ((err2[0] < pos_tol && err2[1] < rot_tol) || it >= max_iters) break;
        used++;
        std::vector<int> dofs;
        for (int d = 0; d < m.nu; d++) if (m.affects[b * m.nv + d]) dofs.push_back(d);
        const int nd = (int)dofs.size();
        std::vector<T> J(6 * nd);
        for (int j = 0; j < nd; j++) {
            T jv[3], jw[3];
            jac_col(m, k, b, dofs[j], xp, jv, jw);
            for (int c = 0; c < 3; c++) { J[c * nd + j] = jv[c]; J[(3 + c) * nd + j] = jw[c]; }
        }
        std::vector<T> A(36, 0);
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) {
                T s = r == c ? damping * damping : (T)0;
                for (int j = 0; j < nd; j++) s += J[r * nd + j] * J[c * nd + j];
                A[r * 6 + c] = s;
            }
        T y[6];
        chol(A, 6);
        chol_solve(A, 6, e, y);
        for (int j = 0; j < nd; j++) {
            T dq = 0;
            for (int r = 0; r < 6; r++) dq += J[r * nd + j] * y[r];
            dq = std::min(std::max(dq, -max_step), max_step);
            const int d = dofs[j];
            T nq = qfull[d] + dq;                               // robot dofs: qadr == dofadr (hinge / slide only)
            if (respect_limits && m.limited[d]) nq = std::min(std::max(nq, m.lim_lo[d]), m.lim_hi[d]);
            qfull[d] = nq;
        }
    }
    return used;
}

// ------------------------------------------------------------------------------------ Philox4x32-10
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// env-parallel loop on std::thread (no OpenMP runtime in this image's default g++ spec path)
template <class F> static void parallel_for(int n, int nthreads, F f) {
    if (nthreads <= 1 || n < 2) { for (int e = 0; e < n; e++) f(e); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    auto work = [&]() { for (;;) { int b = next.fetch_add(16); if (b >= n) break; for (int e = b; e < std::min(n, b + 16); e++) f(e); } };
    for (int t = 0; t < nthreads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

void* orc_create(const OrcModel* c) {
    Handle* h = new Handle();
    build(h->d, c);
    build(h->f, c);
    return h;
}
void orc_destroy(void* h) { delete (Handle*)h; }

void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

// device-reset spawn draw: counter (env_id, episode, 0, 0), key = seed; u = (x>>8)*2^-24; x = fma(u, hi-lo, lo)
int orc_env_dim(const void* hv) { return ((const Handle*)hv)->f.task == 1 ? 14 : 11; }

// object k of a multi-object scene: counter (env_id, episode, k, attempt); object 0 of a one-object scene = orc_spawn_xy
static void spawn_obj_draw(const Model<float>& m, uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t k, uint32_t attempt, float* xy) {
    uint32_t ctr[4] = {env_id, episode, k, attempt}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    philox4x32_10(ctr, key, o);
    float u0 = (float)(o[0] >> 8) * 5.9604644775390625e-8f, u1 = (float)(o[1] >> 8) * 5.9604644775390625e-8f;
    const float* r = &m.obj_spawn[k * 4];
    xy[0] = std::fmaf(u0, r[1] - r[0], r[0]);
    xy[1] = std::fmaf(u1, r[3] - r[2], r[2]);
}
// all objects of one env (xy: n_obj x 2).  With min_dist > 0 (so101/cube_stack_batch.py:70-86) objects 0 and 1 are re-drawn
// together, attempt 1, 2, ... <= 15, until fmaf(dx, dx, dy * dy) >= min_dist^2 in float32.
void orc_spawn_all(const void* hv, uint64_t seed, uint32_t env_id, uint32_t episode, float* xy) {
    const Model<float>& m = ((const Handle*)hv)->f;
    for (int k = 0; k < m.n_obj; k++) spawn_obj_draw(m, seed, env_id, episode, k, 0, xy + 2 * k);
    const float md = m.task_opt[2];
    if (md > 0 && m.n_obj > 1)
        for (uint32_t a = 1; a < 16; a++) {
            const float dx = xy[2] - xy[0], dy = xy[3] - xy[1];
            if (std::fmaf(dx, dx, dy * dy) >= md * md) break;
            spawn_obj_draw(m, seed, env_id, episode, 0, a, xy);
            spawn_obj_draw(m, seed, env_id, episode, 1, a, xy + 2);
        }
}
void orc_spawn_obj(const void* hv, uint64_t seed, uint32_t env_id, uint32_t episode, uint32_t k, float* xy) {
    const Model<float>& m = ((const Handle*)hv)->f;
    float all[2 * 16];
    orc_spawn_all(hv, seed, env_id, episode, all);
    xy[0] = all[2 * k]; xy[1] = all[2 * k + 1];
    (void)m;
}

void orc_spawn_xy(const void* hv, uint64_t seed, uint32_t env_id, uint32_t episode, float* xy) {
    const Model<float>& m = ((const Handle*)hv)->f;
    uint32_t ctr[4] = {env_id, episode, 0, 0}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    philox4x32_10(ctr, key, o);
    float u0 = (float)(o[0] >> 8) * 5.9604644775390625e-8f, u1 = (float)(o[1] >> 8) * 5.9604644775390625e-8f;
    xy[0] = std::fmaf(u0, m.opt[OPT_SPAWN_XHI] - m.opt[OPT_SPAWN_XLO], m.opt[OPT_SPAWN_XLO]);
    xy[1] = std::fmaf(u1, m.opt[OPT_SPAWN_YHI] - m.opt[OPT_SPAWN_YLO], m.opt[OPT_SPAWN_YLO]);
}

#define ORC_API(T, SUF)                                                                                              \
    void orc_fk_##SUF(const void* hv, const T* q, T* xpos, T* xquat) {                                               \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        Kin<T> k; forward_kinematics(m, q, k);                                                                       \
        std::memcpy(xpos, k.xpos.data(), sizeof(T) * m.nb * 3); std::memcpy(xquat, k.xquat.data(), sizeof(T) * m.nb * 4); \
    }                                                                                                                \
    void orc_dynamics_##SUF(const void* hv, const T* q, const T* qd, const T* target_u, T* M, T* Mt, T* bias, T* tau, T* as) { \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        Kin<T> k; Dyn<T> d; forward_kinematics(m, q, k);                                                             \
        std::vector<T> target(m.nv, 0); for (int i = 0; i < m.nu; i++) target[i] = target_u[i];                      \
        dynamics(m, k, q, qd, target.data(), d);                                                                     \
        std::vector<T> L = d.Mt; chol(L, m.nv); chol_solve(L, m.nv, d.fs.data(), d.as.data());                       \
        std::memcpy(M, d.M.data(), sizeof(T) * m.nv * m.nv); std::memcpy(Mt, d.Mt.data(), sizeof(T) * m.nv * m.nv);  \
        std::memcpy(bias, d.bias.data(), sizeof(T) * m.nv); std::memcpy(tau, d.tau.data(), sizeof(T) * m.nv);        \
        std::memcpy(as, d.as.data(), sizeof(T) * m.nv);                                                              \
    }                                                                                                                \
    /* contacts: out[i*9 + (0..2 p, 3..5 n, 6 depth, 7 body a, 8 body b)] */                                         \
    int orc_contacts_##SUF(const void* hv, const T* q, T* out, int cap) {                                            \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        Kin<T> k; std::vector<Contact<T>> cons; forward_kinematics(m, q, k); collide(m, k, cons);                    \
        int n = std::min((int)cons.size(), cap);                                                                     \
        for (int i = 0; i < n; i++) {                                                                                \
            for (int c = 0; c < 3; c++) { out[i * 9 + c] = cons[i].p[c]; out[i * 9 + 3 + c] = cons[i].n[c]; }        \
            out[i * 9 + 6] = cons[i].depth; out[i * 9 + 7] = (T)cons[i].ba; out[i * 9 + 8] = (T)cons[i].bb;          \
        }                                                                                                            \
        return (int)cons.size();                                                                                     \
    }                                                                                                                \
    int orc_rows_##SUF(const void* hv, const T* q, const T* qd, T* J, T* aref, T* D, int cap) {                      \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        Kin<T> k; std::vector<Contact<T>> cons; Rows<T> r; forward_kinematics(m, q, k); collide(m, k, cons);         \
        make_rows(m, k, q, qd, cons, r);                                                                             \
        int n = std::min(r.n, cap);                                                                                  \
        std::memcpy(J, r.J.data(), sizeof(T) * n * m.nv); std::memcpy(aref, r.aref.data(), sizeof(T) * n);           \
        std::memcpy(D, r.D.data(), sizeof(T) * n);                                                                   \
        return r.n;                                                                                                  \
    }                                                                                                                \
    /* state is env-major: q[n][nq], qd[n][nv], qacc[n][nv], target[n][nu]; info[n][4] = ncon,nrows,iters,0 */       \
    void orc_step_##SUF(const void* hv, int n, T* q, T* qd, T* qacc, const T* target, int iterations, double tol,    \
                        int32_t* info, double* grad, int nthreads) {                                                 \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        parallel_for(n, nthreads, [&](int e) {                                                                       \
            StepOut<T> o = substep(m, q + (size_t)e * m.nq, qd + (size_t)e * m.nv, qacc + (size_t)e * m.nv,          \
                                   target + (size_t)e * m.nu, iterations, (T)tol);                                   \
            if (info) { info[e * 4] = o.ncon; info[e * 4 + 1] = o.nrows; info[e * 4 + 2] = o.iters; info[e * 4 + 3] = 0; } \
            if (grad) grad[e] = (double)o.grad;                                                                      \
        });                                                                                                          \
    }                                                                                                                \
    void orc_obs_##SUF(const void* hv, int n, const T* q, T* agent, int agent_dim, T* envs, T* reward, uint8_t* term) { \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        const int ed = m.task == 1 ? 14 : 11;                                                                        \
        for (int e = 0; e < n; e++)                                                                                  \
            observe(m, q + (size_t)e * m.nq, agent + (size_t)e * agent_dim, agent_dim, envs + (size_t)e * ed, reward + e, term + e); \
    }                                                                                                                \
    /* masked reset of the state (no settle step here): object k at (xy[e][k], spawn_z, spawn_quat), robot home, zero vel/qacc */ \
    void orc_reset_##SUF(const void* hv, int n, T* q, T* qd, T* qacc, const uint8_t* mask, const T* xy) {             \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        for (int e = 0; e < n; e++) {                                                                                \
            if (mask && !mask[e]) continue;                                                                          \
            T* qq = q + (size_t)e * m.nq;                                                                            \
            for (int i = 0; i < m.nu; i++) qq[i] = m.home_q[i];                                                      \
            for (int k = 0; k < m.n_obj; k++) {                                                                      \
                T* o = qq + m.nu + 7 * k;                                                                            \
                o[0] = xy[((size_t)e * m.n_obj + k) * 2]; o[1] = xy[((size_t)e * m.n_obj + k) * 2 + 1]; o[2] = m.opt[OPT_SPAWN_Z]; \
                for (int i = 0; i < 4; i++) o[3 + i] = m.spawn_quat[i];                                              \
            }                                                                                                        \
            for (int i = 0; i < m.nv; i++) { qd[(size_t)e * m.nv + i] = 0; qacc[(size_t)e * m.nv + i] = 0; }         \
        }                                                                                                            \
    }                                                                                                                \
    /* batched IK of the eef body: q_init / q_out (n, nu); tpos (n,3) / tquat (n,4) nullable; err (n,2); iters (n) */   \
    /* max_samples > 1: rows that miss the tolerances are re-started SEQUENTIALLY from random arm configurations (uniform  \
       inside the joint limits, Philox counter (row, sample, 'IKRS', dof / 4), key = seed; float32 draw as on the device);      \
       result = first attempt that meets the tolerances, else the attempt with the smallest max(e_pos/pos_tol, e_rot/rot_tol). \
       n_arm = number of leading robot dofs that are re-drawn (the gripper dofs keep their start values).                      */ \
    void orc_ik_##SUF(const void* hv, int n, const T* q_init, const T* tpos, const T* tquat, int max_iters, double damping, \
                      double pos_tol, double rot_tol, double max_step, int respect_limits, const double* mask6, T* q_out, \
                      T* err, int32_t* iters, int max_samples, uint64_t seed, int n_arm) {                          \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        T mk[6]; for (int i = 0; i < 6; i++) mk[i] = (T)mask6[i];                                                    \
        for (int e = 0; e < n; e++) {                                                                                \
            std::vector<T> qf(m.nq, 0), best(m.nu);                                                                  \
            for (int i = 0; i < m.nu; i++) qf[i] = q_init[(size_t)e * m.nu + i];                                     \
            qf[m.nu + 3] = 1;                                                                                        \
            const T* tp_ = tpos ? tpos + (size_t)e * 3 : (const T*)nullptr;                                          \
            const T* tq_ = tquat ? tquat + (size_t)e * 4 : (const T*)nullptr;                                        \
            T e2[2], be[2];                                                                                          \
            int it = inverse_kinematics(m, qf.data(), tp_, tq_, max_iters, (T)damping, (T)pos_tol, (T)rot_tol, (T)max_step, respect_limits, mk, e2); \
            for (int i = 0; i < m.nu; i++) best[i] = qf[i];                                                          \
            be[0] = e2[0]; be[1] = e2[1];                                                                            \
            T bscore = std::max(e2[0] / (T)pos_tol, e2[1] / (T)rot_tol);                                             \
            bool ok = e2[0] < (T)pos_tol && e2[1] < (T)rot_tol;                                                      \
            for (int s = 1; s < max_samples && !ok; s++) {                                                           \
                for (int i = 0; i < m.nu; i++) qf[i] = best[i];                                                      \
                for (int blk = 0; blk < (n_arm + 3) / 4; blk++) {                                                    \
                    uint32_t ctr[4] = {(uint32_t)e, (uint32_t)s, 0x494B5253u, (uint32_t)blk};                       \
                    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];                                \
                    philox4x32_10(ctr, key, o);                                                                      \
                    for (int k = 0; k < 4 && blk * 4 + k < n_arm; k++) {                                             \
                        const int i = blk * 4 + k;                                                                   \
                        const float u = (float)(o[k] >> 8) * 5.9604644775390625e-8f;                                 \
                        qf[i] = (T)std::fmaf(u, (float)m.lim_hi[i] - (float)m.lim_lo[i], (float)m.lim_lo[i]);        \
                    }                                                                                                \
                }                                                                                                    \
                it += inverse_kinematics(m, qf.data(), tp_, tq_, max_iters, (T)damping, (T)pos_tol, (T)rot_tol, (T)max_step, respect_limits, mk, e2); \
                ok = e2[0] < (T)pos_tol && e2[1] < (T)rot_tol;                                                       \
                const T sc = std::max(e2[0] / (T)pos_tol, e2[1] / (T)rot_tol);                                       \
                if (ok || sc < bscore) { bscore = sc; be[0] = e2[0]; be[1] = e2[1]; for (int i = 0; i < m.nu; i++) best[i] = qf[i]; } \
            }                                                                                                        \
            for (int i = 0; i < m.nu; i++) q_out[(size_t)e * m.nu + i] = best[i];                                    \
            if (err) { err[e * 2] = be[0]; err[e * 2 + 1] = be[1]; }                                                 \
            if (iters) iters[e] = it;                                                                                \
        }                                                                                                            \
    }                                                                                                                \
    /* inverse weights recomputed from scratch at q0 (robot dofs) -- independent of the host model compiler */       \
    void orc_invweight_##SUF(const void* hv, const T* q0, T* body_w, T* dof_w) {                                     \
        const Model<T>& m = Pick<T>::get((const Handle*)hv);                                                         \
        Kin<T> k; Dyn<T> d; forward_kinematics(m, q0, k);                                                            \
        std::vector<T> zero(m.nv, 0); dynamics(m, k, q0, zero.data(), zero.data(), d);                               \
        std::vector<T> L = d.M; chol(L, m.nv);                                                                       \
        std::vector<T> col(m.nv), x(m.nv);                                                                           \
        for (int i = 0; i < m.nv; i++) { std::fill(col.begin(), col.end(), (T)0); col[i] = 1; chol_solve(L, m.nv, col.data(), x.data()); dof_w[i] = x[i]; } \
        for (int b = 0; b < m.nb; b++) {                                                                             \
            T tr = 0;                                                                                                \
            for (int r = 0; r < 3; r++) {                                                                            \
                for (int dd = 0; dd < m.nv; dd++) { T jv[3]; jac_col(m, k, b, dd, &k.cpos[b * 3], jv, (T*)nullptr); col[dd] = jv[r]; } \
                chol_solve(L, m.nv, col.data(), x.data());                                                           \
                for (int dd = 0; dd < m.nv; dd++) tr += col[dd] * x[dd];                                             \
            }                                                                                                        \
            body_w[b] = tr / 3;                                                                                      \
        }                                                                                                            \
    }

ORC_API(double, f64)
ORC_API(float, f32)

int orc_num_threads(void) { int n = (int)std::thread::hardware_concurrency(); return n > 0 ? n : 1; }

}  // extern "C"
