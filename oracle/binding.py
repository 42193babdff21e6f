"""ctypes loader for the CPU oracle (`oracle/liboracle.so`).

TEST INFRASTRUCTURE ONLY (parity unpinned, see oracle/oracle.cpp header).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_INT_FIELDS = ["parent", "jtype", "dofadr", "qadr", "limited", "ctrl", "geom_body", "geom_group", "obj_body"]
_DBL_FIELDS = ["pos", "quat", "axis", "mass", "com", "inertia", "body_invweight", "dof_invweight",
               "armature", "damping", "kp", "kv", "frc_lo", "frc_hi", "lim_lo", "lim_hi",
               "geom_half", "geom_pos", "geom_quat", "opt", "spawn_quat", "home_q", "obj_spawn", "task_opt"]
_SCALARS = ["nb", "nv", "nq", "nu", "ngeom", "eef_body", "cube_body", "max_contacts", "n_obj", "task"]


class _OrcModel(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in _SCALARS]
                + [(n, C.POINTER(C.c_int32)) for n in _INT_FIELDS]
                + [(n, C.POINTER(C.c_double)) for n in _DBL_FIELDS])


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.POINTER(_OrcModel)]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Generic-tree CPU oracle bound to one `SceneModel` (gym_genesis_b200.model)."""

    def __init__(self, scene_model, dtype=np.float64):
        self.model = scene_model
        self.flat = scene_model.flat()
        self.dtype = np.dtype(dtype)
        self.suf = "f64" if self.dtype == np.float64 else "f32"
        self.ct = C.c_double if self.dtype == np.float64 else C.c_float
        cm = _OrcModel()
        self._keep = []
        for n in _SCALARS:
            setattr(cm, n, int(self.flat[n]))
        for n in _INT_FIELDS:
            a = np.ascontiguousarray(self.flat[n], dtype=np.int32)
            self._keep.append(a)
            setattr(cm, n, a.ctypes.data_as(C.POINTER(C.c_int32)))
        for n in _DBL_FIELDS:
            a = np.ascontiguousarray(self.flat[n], dtype=np.float64)
            self._keep.append(a)
            setattr(cm, n, a.ctypes.data_as(C.POINTER(C.c_double)))
        self.h = C.c_void_p(lib().orc_create(C.byref(cm)))
        for k in ("nb", "nv", "nq", "nu"):
            setattr(self, k, int(self.flat[k]))
        self.agent_dim = scene_model.nu if getattr(scene_model, "agent_mode", "eef") == "qpos" else 7 + scene_model.n_fing
        self.n_obj = int(self.flat["n_obj"])
        self.env_dim = 14 if scene_model.task == "stack" else 11

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def _fn(self, name):
        return getattr(lib(), f"orc_{name}_{self.suf}")

    def _a(self, x, shape=None):
        a = np.ascontiguousarray(x, dtype=self.dtype)
        if shape is not None:
            a = a.reshape(shape)
        return a

    # ---- diagnostics -------------------------------------------------------------------------
    def fk(self, q):
        q = self._a(q)
        xpos = np.zeros((self.nb, 3), self.dtype)
        xquat = np.zeros((self.nb, 4), self.dtype)
        self._fn("fk")(self.h, _p(q), _p(xpos), _p(xquat))
        return xpos, xquat

    def dynamics(self, q, qd, target):
        q, qd, target = self._a(q), self._a(qd), self._a(target)
        nv = self.nv
        M, Mt = np.zeros((nv, nv), self.dtype), np.zeros((nv, nv), self.dtype)
        bias, tau, a_s = (np.zeros(nv, self.dtype) for _ in range(3))
        self._fn("dynamics")(self.h, _p(q), _p(qd), _p(target), _p(M), _p(Mt), _p(bias), _p(tau), _p(a_s))
        return dict(M=M, Mt=Mt, bias=bias, tau=tau, a_smooth=a_s)

    def contacts(self, q, cap=128):
        q = self._a(q)
        out = np.zeros((cap, 9), self.dtype)
        f = self._fn("contacts")
        f.restype = C.c_int
        n = f(self.h, _p(q), _p(out), C.c_int(cap))
        return out[: min(n, cap)]

    def rows(self, q, qd, cap=512):
        q, qd = self._a(q), self._a(qd)
        J = np.zeros((cap, self.nv), self.dtype)
        aref, D = np.zeros(cap, self.dtype), np.zeros(cap, self.dtype)
        f = self._fn("rows")
        f.restype = C.c_int
        n = f(self.h, _p(q), _p(qd), _p(J), _p(aref), _p(D), C.c_int(cap))
        n = min(n, cap)
        return J[:n], aref[:n], D[:n]

    def invweight(self, q0_full):
        q0 = self._a(q0_full)
        bw, dw = np.zeros(self.nb, self.dtype), np.zeros(self.nv, self.dtype)
        self._fn("invweight")(self.h, _p(q0), _p(bw), _p(dw))
        return bw, dw

    # ---- batched env-major API ---------------------------------------------------------------
    def step(self, q, qd, qacc, target, iterations=100, tol=1e-12, nthreads=1):
        """In-place substep of n envs. q (n,nq), qd (n,nv), qacc (n,nv) warm start, target (n,nu)."""
        n = q.shape[0]
        assert q.dtype == self.dtype and q.flags.c_contiguous and qd.flags.c_contiguous and qacc.flags.c_contiguous
        target = self._a(target, (n, self.nu))
        info = np.zeros((n, 4), np.int32)
        grad = np.zeros(n, np.float64)
        self._fn("step")(self.h, C.c_int(n), _p(q), _p(qd), _p(qacc), _p(target), C.c_int(iterations),
                         C.c_double(tol), _p(info), _p(grad), C.c_int(nthreads))
        return info, grad

    def obs(self, q):
        q = self._a(q)
        n = q.shape[0]
        agent = np.zeros((n, self.agent_dim), self.dtype)
        envs = np.zeros((n, self.env_dim), self.dtype)
        reward = np.zeros(n, self.dtype)
        term = np.zeros(n, np.uint8)
        self._fn("obs")(self.h, C.c_int(n), _p(q), _p(agent), C.c_int(self.agent_dim), _p(envs), _p(reward), _p(term))
        return agent, envs, reward, term.astype(bool)

    def reset(self, q, qd, qacc, xy, mask=None):
        n = q.shape[0]
        xy = self._a(xy, (n, self.n_obj, 2))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._fn("reset")(self.h, C.c_int(n), _p(q), _p(qd), _p(qacc), _p(m) if m is not None else None, _p(xy))

    def ik(self, q_init, target_pos=None, target_quat=None, max_iters=20, damping=0.01, pos_tol=5e-4, rot_tol=5e-3,
           max_step=0.5, respect_limits=True, mask6=None, max_samples=1, seed=0):
        """Batched damped-least-squares IK of the eef body (restates robot.inverse_kinematics, reference
        examples/franka/pick_cube_state.py:46-51).  q_init (n, nu) -> (q (n, nu), err (n, 2), iters (n))."""
        q_init = self._a(q_init)
        n = q_init.shape[0]
        tp = None if target_pos is None else self._a(target_pos, (n, 3))
        tq = None if target_quat is None else self._a(target_quat, (n, 4))
        mk = np.ascontiguousarray(np.ones(6) if mask6 is None else mask6, dtype=np.float64)
        q_out = np.zeros((n, self.nu), self.dtype)
        err = np.zeros((n, 2), self.dtype)
        iters = np.zeros(n, np.int32)
        self._fn("ik")(self.h, C.c_int(n), _p(q_init), _p(tp) if tp is not None else None, _p(tq) if tq is not None else None,
                       C.c_int(max_iters), C.c_double(damping), C.c_double(pos_tol), C.c_double(rot_tol), C.c_double(max_step),
                       C.c_int(int(respect_limits)), _p(mk), _p(q_out), _p(err), _p(iters), C.c_int(int(max_samples)),
                       C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_int(self.model.n_arm))
        return q_out, err, iters

    def new_state(self, n):
        return (np.zeros((n, self.nq), self.dtype), np.zeros((n, self.nv), self.dtype), np.zeros((n, self.nv), self.dtype))

    def spawn_xy(self, seed, env_id, episode):
        out = np.zeros(2, np.float32)
        lib().orc_spawn_xy(self.h, C.c_uint64(seed), C.c_uint32(env_id), C.c_uint32(episode), _p(out))
        return out


def _spawn_obj(self, seed, env_id, episode, k):
    out = np.zeros(2, np.float32)
    lib().orc_spawn_obj(self.h, C.c_uint64(seed), C.c_uint32(env_id), C.c_uint32(episode), C.c_uint32(k), _p(out))
    return out


Oracle.spawn_obj = _spawn_obj


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, np.uint32)
    lib().orc_philox(_p(c), _p(k), _p(o))
    return o


def num_threads():
    return int(lib().orc_num_threads())


# ---- timing builds (bench.py cpu_baseline / --impl reference legs only) ------------------------------------------------------------
_fast_built = False


def build_fast() -> None:
    """(Re)compile the -O3 -march=native timing libraries ON THIS MACHINE (the in-tree copies may come from another CPU)."""
    global _fast_built
    if not _fast_built:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_fast.so", "libhostport.so"])
        _fast_built = True


class FastOracle(Oracle):
    """The generic oracle built with -O3 -march=native (same source; contraction allowed): timing only, never the checker."""

    def _fn(self, name):
        if not hasattr(FastOracle, "_flib"):
            build_fast()
            FastOracle._flib = C.CDLL(os.path.join(_HERE, "liboracle_fast.so"))
            FastOracle._flib.orc_create.restype = C.c_void_p
            FastOracle._flib.orc_create.argtypes = [C.POINTER(_OrcModel)]
        return getattr(FastOracle._flib, f"orc_{name}_{self.suf}")


class HostPort:
    """The specialised per-env pipeline (gym_genesis_b200/csrc/gg_env.cuh compiled for the host, oracle/hostport.cpp) on host threads:
    one call = one workload step (masked reset + substep + observation + flags) of n envs, float32, the GPU's solver settings."""

    def __init__(self, scene_model):
        from gym_genesis_b200 import cabi
        build_fast()
        self.lib = C.CDLL(os.path.join(_HERE, "libhostport.so"))
        self.gg = cabi.compile_gg(scene_model)
        assert self.lib.hp_sizeof_model() == C.sizeof(self.gg)
        self.nu = scene_model.nu
        self.agent_dim = 7 + scene_model.n_fing

    def new_state(self, n):
        return (np.zeros((n, self.nu + 7), np.float32), np.zeros((n, self.nu + 6), np.float32), np.zeros((n, self.nu + 6), np.float32))

    def step(self, q, qd, qacc, target, reset_mask=None, xy=None, iterations=10, tol=1e-5, nthreads=1):
        n = q.shape[0]
        if not hasattr(self, "_out") or self._out[0].shape[0] != n:
            self._out = (np.zeros((n, self.agent_dim), np.float32), np.zeros((n, 11), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8))
        a, e, r, t = self._out
        target = np.ascontiguousarray(target, np.float32)
        m = None if reset_mask is None else np.ascontiguousarray(reset_mask, np.uint8)
        x = None if xy is None else np.ascontiguousarray(xy, np.float32)
        self.lib.hp_step(C.byref(self.gg), C.c_int(n), _p(q), _p(qd), _p(qacc), _p(target), _p(m) if m is not None else None,
                         _p(x) if x is not None else None, C.c_int(iterations), C.c_float(tol), _p(a), _p(e), _p(r), _p(t), C.c_int(nthreads))
        return a, e, r, t

    def run_workload(self, q, qd, qacc, pool, steps, reset_prob, seed=0, iterations=10, tol=1e-5, nthreads=1):
        """`steps` workload steps entirely in native code (masked resets + substep + observation per env and step), threads own env
        ranges for the whole run: the CPU baseline bench.py quotes.  pool: (npool, n, nu) float32 action batches, cycled."""
        n = q.shape[0]
        pool = np.ascontiguousarray(pool, np.float32)
        a, e = np.zeros((n, self.agent_dim), np.float32), np.zeros((n, 11), np.float32)
        r, t = np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self.lib.hp_run_workload(C.byref(self.gg), C.c_int(n), C.c_int(steps), C.c_int(nthreads), _p(q), _p(qd), _p(qacc), _p(pool),
                                 C.c_int(pool.shape[0]), C.c_float(reset_prob), C.c_uint64(seed), C.c_int(iterations), C.c_float(tol),
                                 _p(a), _p(e), _p(r), _p(t))
        return a, e, r, t
