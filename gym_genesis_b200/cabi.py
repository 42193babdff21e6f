"""ctypes mirror of `include/gg_b200.h` and the host-side model compiler.

The product path is the CUDA library `libgg_b200.so` built in-tree by `__graft_entry__.build()`.
There is deliberately NO fallback: if the library is missing or fails to load, importing the
simulator raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import model as mdl

GG_MAX_ARM, GG_MAX_FING, GG_MAX_DOF, GG_MAX_PADS, GG_MAX_CONTACTS = 7, 2, 9, 12, 64
GG_MAX_OBJ, GG_STACK_MAX_CONTACTS, GG_STACK_ENV_STATE_DIM = 5, 96, 14

_f = C.c_float
_i = C.c_int32


class GgModel(C.Structure):
    _fields_ = [
        ("n_arm", _i), ("n_fing", _i), ("fing_hinge", _i), ("n_pads", _i),
        ("settle_steps", _i), ("horizon", _i), ("max_contacts", _i), ("reserved0", _i),
        ("base_pos", _f * 3), ("base_quat", _f * 4),
        ("link_pos", (_f * 3) * GG_MAX_ARM), ("link_quat", (_f * 4) * GG_MAX_ARM),
        ("link_mass", _f * GG_MAX_ARM), ("link_com", (_f * 3) * GG_MAX_ARM), ("link_inertia", (_f * 6) * GG_MAX_ARM),
        ("fing_pos", (_f * 3) * GG_MAX_FING), ("fing_quat", (_f * 4) * GG_MAX_FING), ("fing_axis", (_f * 3) * GG_MAX_FING),
        ("fing_mass", _f * GG_MAX_FING), ("fing_com", (_f * 3) * GG_MAX_FING), ("fing_inertia", (_f * 6) * GG_MAX_FING),
        ("eef_pos", _f * 3), ("eef_quat", _f * 4),
        ("armature", _f * GG_MAX_DOF), ("damping", _f * GG_MAX_DOF), ("kp", _f * GG_MAX_DOF), ("kv", _f * GG_MAX_DOF),
        ("frc_lo", _f * GG_MAX_DOF), ("frc_hi", _f * GG_MAX_DOF), ("lim_lo", _f * GG_MAX_DOF), ("lim_hi", _f * GG_MAX_DOF),
        ("dof_invweight", _f * GG_MAX_DOF),
        ("body_invweight", _f * (1 + GG_MAX_FING)),
        ("pad_body", _i * GG_MAX_PADS),
        ("pad_half", (_f * 3) * GG_MAX_PADS), ("pad_pos", (_f * 3) * GG_MAX_PADS), ("pad_quat", (_f * 4) * GG_MAX_PADS),
        ("pad_bound", (_f * 4) * (1 + GG_MAX_FING)),
        ("cube_half", _f * 3), ("cube_mass", _f), ("cube_inertia", _f * 3), ("cube_invweight", _f),
        ("dt", _f), ("gravity", _f), ("plane_z", _f), ("friction", _f),
        ("tc", _f), ("dr", _f), ("dmin", _f), ("dmax", _f), ("width", _f), ("mid", _f), ("power", _f), ("minval", _f),
        ("home_q", _f * GG_MAX_DOF),
        ("spawn_lo", _f * 2), ("spawn_hi", _f * 2), ("spawn_z", _f), ("spawn_quat", _f * 4), ("success_z", _f),
        ("reg_scale", _f), ("reserved1", _f * 3),
    ]


class GgStackTask(C.Structure):
    _fields_ = [("n_obj", _i), ("max_contacts", _i),
                ("spawn_lo", (_f * 2) * GG_MAX_OBJ), ("spawn_hi", (_f * 2) * GG_MAX_OBJ),
                ("stack_xy", _f), ("stack_z", _f), ("min_dist", _f), ("agent_qpos", _i)]


class GgConfig(C.Structure):
    _fields_ = [
        ("solver_iterations", _i), ("solver_tolerance", _f), ("auto_reset", _i), ("time_limit", _i),
        ("reset_prob", _f), ("seed", C.c_uint64), ("first_env_id", C.c_int64),
    ]


class GgHostIO(C.Structure):
    _fields_ = [("action", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p), ("truncated", C.c_void_p),
                ("wait", _i), ("reserved0", _i), ("obs_agent", C.c_void_p), ("obs_env", C.c_void_p), ("action_dev", C.c_void_p)]


class GgIkConfig(C.Structure):
    _fields_ = [("max_iters", _i), ("damping", _f), ("pos_tol", _f), ("rot_tol", _f), ("max_step", _f), ("respect_limits", _i),
                ("pos_mask", _f * 3), ("rot_mask", _f * 3), ("max_samples", _i), ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32)]


def _set(arr, values):
    values = np.asarray(values, dtype=np.float64).ravel()
    for k, v in enumerate(values):
        arr[k] = float(v)


def compile_gg(sm: "mdl.SceneModel") -> GgModel:
    """Compile a generic `SceneModel` tree into the specialised constants the CUDA kernels read:
    a serial chain of `n_arm` z-hinges, fixed children merged into their arm link, gripper dofs as
    children of the last arm link, pads expressed in their (merged) body frame."""
    B = sm.bodies
    arm = [i for i, b in enumerate(B) if b.jtype == mdl.HINGE and sm.dofadr[i] < sm.n_arm]
    fing = [i for i, b in enumerate(B) if b.jtype in (mdl.HINGE, mdl.SLIDE) and sm.n_arm <= sm.dofadr[i] < sm.nu]
    assert len(arm) == sm.n_arm and len(fing) == sm.n_fing
    assert sm.n_arm <= GG_MAX_ARM and sm.n_fing <= GG_MAX_FING
    if mdl.GENESIS_COMPAT_POWER != 2.0:
        raise ValueError("the CUDA impedance code is specialised for solimp power = 2")

    def rel_transform(child, ancestor):
        """(R, p) of body `child`'s frame (joint at q=0) expressed in body `ancestor`'s frame; ancestor=-1 -> world."""
        R, p, qq = np.eye(3), np.zeros(3), np.array([1.0, 0.0, 0.0, 0.0])
        i = child
        while i != ancestor:
            b = B[i]
            Rb = mdl.quat_to_mat(b.quat)
            p = np.asarray(b.pos, dtype=np.float64) + Rb @ p
            R = Rb @ R
            qq = mdl.quat_mul(mdl.quat_normalize(b.quat), qq)     # quaternion chain keeps the MJCF sign convention
            i = b.parent
            assert i >= -1
            if i != ancestor and i >= 0:
                assert B[i].jtype == mdl.FIXED, "only fixed bodies may sit between chain links"
        rel_transform.quat = qq
        return R, p

    def owner(i):
        """nearest moving ancestor-or-self of body i (-1 = welded to world)."""
        while i >= 0 and B[i].jtype == mdl.FIXED:
            i = B[i].parent
        return i

    g = GgModel()
    g.n_arm, g.n_fing, g.n_pads = sm.n_arm, sm.n_fing, 0
    g.fing_hinge = int(any(B[i].jtype == mdl.HINGE for i in fing))
    g.settle_steps, g.horizon, g.max_contacts = sm.settle_steps, sm.horizon, min(mdl.GENESIS_COMPAT_MAX_CONTACTS, GG_MAX_CONTACTS)
    # base: transform of the first arm link's parent chain up to the world is folded into link 0
    _set(g.base_pos, [0, 0, 0]); _set(g.base_quat, [1, 0, 0, 0])
    prev = -1
    for l, i in enumerate(arm):
        assert np.allclose(B[i].axis, (0, 0, 1)), "arm hinges must rotate about their local z axis"
        R, p = rel_transform(i, prev)
        _set(g.link_pos[l], p); _set(g.link_quat[l], rel_transform.quat)
        prev = i
    # merge fixed descendants into their owning arm link / finger
    def merged_inertia(root):
        parts = []
        for j, b in enumerate(B):
            if b.mass > 0 and owner(j) == root:
                R, p = (np.eye(3), np.zeros(3)) if j == root else rel_transform(j, root)
                parts.append((b.mass, p + R @ np.asarray(b.com, dtype=np.float64), R @ mdl.inertia6_to_mat(b.inertia) @ R.T))
        m = sum(x[0] for x in parts)
        c = sum(x[0] * x[1] for x in parts) / m
        I = np.zeros((3, 3))
        for mi, ci, Ii in parts:
            d = ci - c
            I += Ii + mi * (d @ d * np.eye(3) - np.outer(d, d))
        return m, c, mdl.mat_to_inertia6(I)

    for l, i in enumerate(arm):
        m, c, I6 = merged_inertia(i)
        g.link_mass[l] = m; _set(g.link_com[l], c); _set(g.link_inertia[l], I6)
    last = arm[-1]
    for f, i in enumerate(fing):
        assert owner(B[i].parent) == last, "gripper dofs must hang off the last arm link"
        R, p = rel_transform(i, last)
        _set(g.fing_pos[f], p); _set(g.fing_quat[f], rel_transform.quat); _set(g.fing_axis[f], B[i].axis)
        if B[i].jtype == mdl.HINGE:
            assert np.allclose(B[i].axis, (0, 0, 1))
        m, c, I6 = merged_inertia(i)
        g.fing_mass[f] = m; _set(g.fing_com[f], c); _set(g.fing_inertia[f], I6)
    if sm.eef_body == last:
        pe, qe = np.zeros(3), np.array([1.0, 0.0, 0.0, 0.0])
    else:
        assert owner(sm.eef_body) == last
        _, pe = rel_transform(sm.eef_body, last)
        qe = rel_transform.quat
    _set(g.eef_pos, pe); _set(g.eef_quat, qe)
    dt = sm.dof_table()
    body_w, dof_w = sm.invweight0()
    for name in ("armature", "damping", "kp", "kv", "frc_lo", "frc_hi", "lim_lo", "lim_hi"):
        _set(getattr(g, name), dt[name][: sm.nu])
    assert all(dt["limited"][: sm.nu]), "every robot dof is expected to be range-limited"
    _set(g.dof_invweight, dof_w[: sm.nu])
    _set(g.body_invweight, [body_w[last]] + [body_w[i] for i in fing])
    k = 0
    for geom in sm.geoms:
        if geom.group != 0:
            continue
        ob = owner(geom.body)
        code = 0 if ob == last else 1 + fing.index(ob)
        if ob == last and geom.body != last:
            raise NotImplementedError("pads on fixed children of the last arm link")
        _set(g.pad_half[k], geom.half); _set(g.pad_pos[k], geom.pos); _set(g.pad_quat[k], mdl.quat_normalize(geom.quat))
        g.pad_body[k] = code
        k += 1
    g.n_pads = k
    # conservative bounding sphere of all pads per carrying body (body-level broad phase in the kernel)
    for b in range(1 + GG_MAX_FING):
        corners = []
        for j in range(k):
            if g.pad_body[j] != b:
                continue
            R = mdl.quat_to_mat(list(g.pad_quat[j]))
            for sx in (-1, 1):
                for sy in (-1, 1):
                    for sz in (-1, 1):
                        corners.append(np.array(list(g.pad_pos[j])) + R @ (np.array(list(g.pad_half[j])) * [sx, sy, sz]))
        if corners:
            corners = np.array(corners)
            ctr = 0.5 * (corners.min(0) + corners.max(0))
            rad = np.linalg.norm(corners - ctr, axis=1).max() * 1.001 + 1e-6
            _set(g.pad_bound[b], [*ctr, rad])
        else:
            _set(g.pad_bound[b], [0, 0, 0, -1])
    cube = B[sm.cube_body]
    cg = [x for x in sm.geoms if x.group == 1]
    assert len(cg) == sm.n_obj and [x.body for x in cg] == list(sm.obj_bodies)
    for x, b in zip(cg, sm.obj_bodies):      # every object is the same box (franka/cube_stack.py:54-77: five 0.04 m cubes)
        assert np.allclose(x.half, cg[0].half) and np.allclose(x.pos, 0) and B[b].mass == cube.mass
        assert np.allclose(B[b].inertia, cube.inertia) and abs(body_w[b] - body_w[sm.cube_body]) < 1e-9 * body_w[b]
    _set(g.cube_half, cg[0].half)
    g.cube_mass = cube.mass
    _set(g.cube_inertia, cube.inertia[:3])
    g.cube_invweight = body_w[sm.cube_body]
    g.dt, g.gravity, g.plane_z, g.friction = sm.dt, sm.gravity, sm.plane_z, sm.friction
    g.tc, g.dr, g.dmin, g.dmax = mdl.GENESIS_COMPAT_TIMECONST, mdl.GENESIS_COMPAT_DAMPRATIO, mdl.GENESIS_COMPAT_DMIN, mdl.GENESIS_COMPAT_DMAX
    g.width, g.mid, g.power, g.minval = mdl.GENESIS_COMPAT_WIDTH, mdl.GENESIS_COMPAT_MID, mdl.GENESIS_COMPAT_POWER, mdl.GENESIS_COMPAT_MINVAL
    _set(g.home_q, sm.home_q)
    _set(g.spawn_lo, sm.spawn_lo); _set(g.spawn_hi, sm.spawn_hi); g.spawn_z = sm.spawn_z
    _set(g.spawn_quat, sm.spawn_quat); g.success_z = sm.success_z
    g.reg_scale = sm.reg_scale()
    return g


def compile_stack_task(sm: "mdl.SceneModel") -> GgStackTask:
    """Task constants of a multi-object scene (`franka/cube_stack.py:119-196`): spawn ranges and stacking thresholds."""
    assert sm.task == "stack" and 2 <= sm.n_obj <= GG_MAX_OBJ
    t = GgStackTask()
    t.n_obj, t.max_contacts = sm.n_obj, min(sm.max_contacts, GG_STACK_MAX_CONTACTS)
    for k, (xlo, xhi, ylo, yhi) in enumerate(sm.obj_spawn):
        _set(t.spawn_lo[k], [xlo, ylo]); _set(t.spawn_hi[k], [xhi, yhi])
    t.stack_xy, t.stack_z = sm.stack_xy, sm.stack_z
    t.min_dist, t.agent_qpos = sm.stack_min_dist, int(sm.agent_mode == "qpos")
    return t


# ----------------------------------------------------------------------------------------------
# library loading (no fallback)
# ----------------------------------------------------------------------------------------------
LIB_NAME = "libgg_b200.so"
# GG_B200_LIB: developer override used by tools/ to time alternative builds of the SAME sources (compile-time variants)
LIB_PATH = os.environ.get("GG_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)
EXPORTS = ["gg_last_error", "gg_version", "gg_create", "gg_destroy", "gg_configure", "gg_state_rows", "gg_state_stride",
           "gg_bind_state", "gg_step", "gg_step_host", "gg_reset", "gg_get_state", "gg_set_state", "gg_observe", "gg_stats", "gg_launch_count",
           "gg_ik", "gg_ik_defaults", "gg_frame_dim", "gg_record", "gg_export_successful",
           "gg_stack_create", "gg_stack_destroy", "gg_stack_configure", "gg_stack_state_rows", "gg_stack_state_stride",
           "gg_stack_bind_state", "gg_stack_step", "gg_stack_reset", "gg_stack_get_state", "gg_stack_set_state", "gg_stack_observe",
           "gg_stack_stats", "gg_stack_info", "gg_stack_agent_dim", "gg_stack_frame_dim", "gg_stack_record", "gg_stack_export_successful",
           "gg_stack_step_host"]

_lib = None


def load_library():
    """Load the CUDA C-ABI library.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                          "There is no CPU fallback for the env.step path.")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.gg_last_error.restype = C.c_char_p
    lib.gg_version.restype = C.c_int
    lib.gg_launch_count.restype = C.c_int64
    lib.gg_create.argtypes = [C.POINTER(GgModel), C.c_int, C.c_int, C.POINTER(vp)]
    lib.gg_destroy.argtypes = [vp]
    lib.gg_configure.argtypes = [vp, C.POINTER(GgConfig)]
    lib.gg_state_rows.argtypes = [vp]
    lib.gg_state_stride.argtypes = [vp]
    lib.gg_bind_state.argtypes = [vp, vp, vp, vp]
    lib.gg_step.argtypes = [vp] * 9
    lib.gg_step_host.argtypes = [vp, C.POINTER(GgHostIO)] + [vp] * 7
    lib.gg_reset.argtypes = [vp] * 6
    lib.gg_get_state.argtypes = [vp] * 5
    lib.gg_set_state.argtypes = [vp] * 5
    lib.gg_observe.argtypes = [vp] * 6
    lib.gg_stats.argtypes = [vp, vp, C.c_int, vp]
    lib.gg_frame_dim.argtypes = [vp]
    lib.gg_record.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.gg_export_successful.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp]
    lib.gg_ik_defaults.argtypes = [C.POINTER(GgIkConfig)]
    lib.gg_ik_defaults.restype = None
    lib.gg_ik.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.POINTER(GgIkConfig), vp]
    lib.gg_stack_create.argtypes = [C.POINTER(GgModel), C.POINTER(GgStackTask), C.c_int, C.c_int, C.POINTER(vp)]
    lib.gg_stack_destroy.argtypes = [vp]
    lib.gg_stack_configure.argtypes = [vp, C.POINTER(GgConfig)]
    lib.gg_stack_state_rows.argtypes = [vp]
    lib.gg_stack_state_stride.argtypes = [vp]
    lib.gg_stack_bind_state.argtypes = [vp, vp, vp, vp]
    lib.gg_stack_step.argtypes = [vp] * 9
    lib.gg_stack_reset.argtypes = [vp] * 6
    lib.gg_stack_get_state.argtypes = [vp] * 5
    lib.gg_stack_set_state.argtypes = [vp] * 5
    lib.gg_stack_observe.argtypes = [vp] * 6
    lib.gg_stack_stats.argtypes = [vp, vp, C.c_int, vp]
    lib.gg_stack_info.argtypes = [vp, vp, vp, vp]
    lib.gg_stack_agent_dim.argtypes = [vp]
    lib.gg_stack_step_host.argtypes = [vp, C.POINTER(GgHostIO)] + [vp] * 7
    lib.gg_stack_frame_dim.argtypes = [vp]
    lib.gg_stack_record.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.gg_stack_export_successful.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib
