"""Scene/robot model tables and the host-side model compiler.

The reference builds its scene through Genesis entity calls
(`gym_genesis/tasks/franka/cube_pick.py:34-68`: Plane + MJCF panda.xml + 0.04 m Box,
`SimOptions(dt=0.01)`, `RigidOptions(box_box_detection=True)`).  Neither Genesis nor its
`xml/franka_emika_panda/panda.xml` asset is present in the reference tree, so the Panda
constants below restate the public Franka Emika Panda description;
gains are the ones the reference hard-codes for the same robot
(`gym_genesis/tasks/franka/cube_stack.py:156-161`).

Everything Genesis-specific that is a *recollection* of the upstream solver (SURVEY.md
Appendix A) is a named ``GENESIS_COMPAT_*`` constant here, so a session with Genesis
installed can flip/fix it in one place.  The tables are plain data: the CPU oracle
(`oracle/`) consumes them as a generic kinematic tree, the CUDA path consumes the
*compiled* form produced by :func:`compile_gg` (fixed links merged, rotations pre-multiplied).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

FIXED, HINGE, SLIDE, FREE = 0, 1, 2, 3

# ----------------------------------------------------------------------------------------------
# Genesis-compat constants ([UPSTREAM-RECALL] -> parity unpinned)
# ----------------------------------------------------------------------------------------------
GENESIS_COMPAT_DT = 0.01                 # SimOptions(dt=0.01), franka/cube_pick.py:45; substeps=1
GENESIS_COMPAT_GRAVITY = 9.81            # gravity=(0,0,-9.81)
GENESIS_COMPAT_RIGID_DENSITY = 200.0     # gs.materials.Rigid() default rho -> cube mass 0.0128 kg
GENESIS_COMPAT_FRICTION = 1.0            # default geom friction, contact mu = max(mu_a, mu_b)
GENESIS_COMPAT_TIMECONST = 0.02          # max(constraint_timeconst=0.01, 2*dt)
GENESIS_COMPAT_DAMPRATIO = 1.0
GENESIS_COMPAT_DMIN = 0.9
GENESIS_COMPAT_DMAX = 0.95
GENESIS_COMPAT_WIDTH = 0.001
GENESIS_COMPAT_MID = 0.5
GENESIS_COMPAT_POWER = 2.0
GENESIS_COMPAT_MINVAL = 1e-15            # MuJoCo mjMINVAL floor for the regulariser R
GENESIS_COMPAT_IMPLICIT_KV = True        # approximate_implicitfast: Mt = M + arm + dt*(damping+kv)
GENESIS_COMPAT_FREE_ANG_BODYFRAME = True  # free-joint angular dofs in body frame, quat (x)= exp(w dt)
GENESIS_COMPAT_SELF_COLLISION = False    # pad-pad / pad-link pairs skipped (same entity)
GENESIS_COMPAT_PYRAMID_MU_SCALING = True  # regulariser of a pyramidal contact row R = 2 mu^2 (1 + mu^2) (1-imp)/imp (w_a + w_b) (MuJoCo-lineage);
                                          # False: R = (1-imp)/imp (w_a + w_b).  Open risk: with the SO-101 task's friction 5 the
                                          # scaled form makes the contact 1300x softer than the unscaled one and the resting cube sinks 16 mm
GENESIS_COMPAT_PANDA_WRIST_FORCE = 87.0   # actuator force range of Panda joints 5-7 when the task never calls set_dofs_force_range (CubePick,
                                          # franka/cube_pick.py:86-112): 87 = what the stacking tasks set explicitly (franka/cube_stack.py:158-161);
                                          # the public panda.xml clamps joints 5-7 at +-12 N m (SURVEY.md App. B risk U1) -> pass wrist_force=12
GENESIS_COMPAT_MAX_PLANE_CONTACTS = 4    # mjc_PlaneBox keeps the first 4 penetrating corners
GENESIS_COMPAT_MAX_BOXBOX_CONTACTS = 8
GENESIS_COMPAT_MAX_CONTACTS = 64         # per-env cap (Genesis: max_collision_pairs), deterministic order
GENESIS_COMPAT_MAX_CONTACTS_STACK = 96   # same cap for the five-cube CubeStack scene (20 cube-plane points alone)
# robot.inverse_kinematics defaults (RigidEntity.inverse_kinematics keyword defaults, [UPSTREAM-RECALL, medium])
GENESIS_COMPAT_IK_MAX_ITERS = 20         # max_solver_iters
GENESIS_COMPAT_IK_DAMPING = 0.01         # dq = J^T (J J^T + damping^2 I)^-1 e
GENESIS_COMPAT_IK_POS_TOL = 5e-4         # m
GENESIS_COMPAT_IK_ROT_TOL = 5e-3         # rad
GENESIS_COMPAT_IK_MAX_STEP = 0.5         # max_step_size per dof and update
GENESIS_COMPAT_IK_MAX_SAMPLES = 50       # Genesis retries unsolved rows from random qpos (max_samples, host RNG stream); restated with the
                                         # device Philox stream (same distribution: uniform inside the joint limits), gg_ik.cuh


# ----------------------------------------------------------------------------------------------
# small quaternion helpers (wxyz)
# ----------------------------------------------------------------------------------------------
def quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def quat_to_mat(q):
    w, x, y, z = quat_normalize(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def mat_to_quat(R):
    """Rotation matrix -> unit quaternion wxyz (Shepperd)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def axis_angle_mat(axis, ang):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def inertia6_to_mat(i6):
    xx, yy, zz, xy, xz, yz = i6
    return np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]], dtype=np.float64)


def mat_to_inertia6(I):
    return np.array([I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]])


# ----------------------------------------------------------------------------------------------
# model containers
# ----------------------------------------------------------------------------------------------
@dataclass
class Body:
    name: str
    parent: int                      # index into SceneModel.bodies, -1 = world
    pos: Sequence[float]
    quat: Sequence[float]            # wxyz (normalised on use)
    jtype: int = FIXED
    axis: Sequence[float] = (0.0, 0.0, 1.0)
    range: Sequence[float] = (0.0, 0.0)
    limited: bool = False
    armature: float = 0.0
    damping: float = 0.0
    kp: float = 0.0
    kv: float = 0.0
    frc: Sequence[float] = (0.0, 0.0)
    mass: float = 0.0
    com: Sequence[float] = (0.0, 0.0, 0.0)
    inertia: Sequence[float] = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)   # xx yy zz xy xz yz at COM, body frame


@dataclass
class BoxGeom:
    body: int
    half: Sequence[float]
    pos: Sequence[float] = (0.0, 0.0, 0.0)
    quat: Sequence[float] = (1.0, 0.0, 0.0, 0.0)
    group: int = 0                   # 0 = robot pad, 1 = task object (cube)


@dataclass
class SceneModel:
    """One env's world: a fixed-base arm with a gripper, one free box, a ground plane."""
    name: str
    bodies: List[Body]
    geoms: List[BoxGeom]
    eef_body: int
    cube_body: int
    home_q: np.ndarray               # robot dofs only
    n_arm: int                       # serial hinge chain length (7 Panda, 5 SO-101)
    n_fing: int                      # gripper dofs (2 slides Panda, 1 hinge SO-101)
    plane_z: float = 0.0
    friction: float = GENESIS_COMPAT_FRICTION
    dt: float = GENESIS_COMPAT_DT
    gravity: float = GENESIS_COMPAT_GRAVITY
    # task (reference: franka/cube_pick.py:86-112,130-135)
    spawn_lo: Sequence[float] = (0.45, -0.25)
    spawn_hi: Sequence[float] = (0.80, 0.25)
    spawn_z: float = 0.02
    spawn_quat: Sequence[float] = (0.0, 0.0, 0.0, 1.0)   # literal [0,0,0,1] under wxyz
    success_z: float = 0.1
    settle_steps: int = 1
    horizon: int = 200
    qpos0: Optional[np.ndarray] = None   # configuration at which invweight0 is evaluated (MJCF qpos0 = 0)
    # multi-object scenes (reference: franka/cube_stack.py:27-240): free boxes in dof order, cube_body = obj_bodies[0]
    obj_bodies: Optional[List[int]] = None
    obj_spawn: Optional[Sequence[Sequence[float]]] = None   # per object (x_lo, x_hi, y_lo, y_hi); draw order x then y per object
    task: str = "pick"               # "pick": reward = cube z > success_z; "stack": xy_dist < stack_xy and z_diff > stack_z
    stack_xy: float = 0.05           # franka/cube_stack.py:195
    stack_z: float = 0.03            # franka/cube_stack.py:195
    stack_min_dist: float = 0.0      # so101/cube_stack_batch.py:70 (0.06): cube_1 / cube_2 re-drawn until this far apart
    agent_mode: str = "eef"          # "eef": eef pose + gripper dofs; "qpos": joint positions (so101/cube_stack_batch.py:176)
    max_contacts: int = GENESIS_COMPAT_MAX_CONTACTS
    pyramid_mu_scaling: bool = GENESIS_COMPAT_PYRAMID_MU_SCALING

    def reg_scale(self) -> float:
        """Factor of the pyramidal-row regulariser (see GENESIS_COMPAT_PYRAMID_MU_SCALING)."""
        mu = float(np.float32(self.friction))
        return float(np.float32(2.0 * mu * mu * (1.0 + mu * mu))) if self.pyramid_mu_scaling else 1.0

    # ---- derived sizes -----------------------------------------------------------------------
    def __post_init__(self):
        self.nb = len(self.bodies)
        self.dofadr = np.full(self.nb, -1, dtype=np.int32)
        self.qadr = np.full(self.nb, -1, dtype=np.int32)
        nv = nq = 0
        for i, b in enumerate(self.bodies):
            if b.jtype in (HINGE, SLIDE):
                self.dofadr[i], self.qadr[i] = nv, nq
                nv += 1
                nq += 1
            elif b.jtype == FREE:
                self.dofadr[i], self.qadr[i] = nv, nq
                nv += 6
                nq += 7
        self.nv, self.nq = nv, nq
        self.nu = self.n_arm + self.n_fing
        assert self.dofadr[self.cube_body] == self.nu, "robot dofs must precede the cube's free joint"
        if self.qpos0 is None:
            self.qpos0 = np.zeros(self.nu)
        if self.obj_bodies is None:
            self.obj_bodies = [self.cube_body]
        if self.obj_spawn is None:
            self.obj_spawn = [(self.spawn_lo[0], self.spawn_hi[0], self.spawn_lo[1], self.spawn_hi[1])] * len(self.obj_bodies)
        self.n_obj = len(self.obj_bodies)
        assert self.obj_bodies[0] == self.cube_body and len(self.obj_spawn) == self.n_obj
        for k, b in enumerate(self.obj_bodies):
            assert self.bodies[b].jtype == FREE and self.dofadr[b] == self.nu + 6 * k, "objects must follow the robot in dof order"
        assert self.nv == self.nu + 6 * self.n_obj

    # ---- per-dof tables ----------------------------------------------------------------------
    def dof_table(self):
        """Arrays of length nv: armature, damping, kp, kv, frc_lo, frc_hi, lim_lo, lim_hi, limited, ctrl."""
        nv = self.nv
        t = {k: np.zeros(nv) for k in ("armature", "damping", "kp", "kv", "frc_lo", "frc_hi", "lim_lo", "lim_hi")}
        t["limited"] = np.zeros(nv, dtype=np.int32)
        t["ctrl"] = np.zeros(nv, dtype=np.int32)
        for i, b in enumerate(self.bodies):
            if b.jtype in (HINGE, SLIDE):
                d = self.dofadr[i]
                t["armature"][d], t["damping"][d], t["kp"][d], t["kv"][d] = b.armature, b.damping, b.kp, b.kv
                t["frc_lo"][d], t["frc_hi"][d] = b.frc
                t["lim_lo"][d], t["lim_hi"][d] = b.range
                t["limited"][d] = int(b.limited)
                t["ctrl"][d] = 1
        return t

    # ---- fp64 numpy kinematics / dynamics (host logic; used for invweight0 and golden vectors) -
    def default_cube_q(self):
        if self.n_obj == 1:
            return np.array([0.65, 0.0, self.plane_z + 0.02, 1.0, 0.0, 0.0, 0.0])
        return np.concatenate([[*self.bodies[b].pos, 1.0, 0.0, 0.0, 0.0] for b in self.obj_bodies]).astype(np.float64)

    def full_q(self, q_robot, cube_q=None):
        q = np.zeros(self.nq)
        q[: self.nu] = q_robot
        q[self.nu:] = self.default_cube_q() if cube_q is None else cube_q
        return q

    def fk(self, q):
        """World pose of every body: (xpos (nb,3), xmat (nb,3,3))."""
        xpos = np.zeros((self.nb, 3))
        xmat = np.zeros((self.nb, 3, 3))
        for i, b in enumerate(self.bodies):
            if b.jtype == FREE:
                a = self.qadr[i]
                xpos[i] = q[a:a + 3]
                xmat[i] = quat_to_mat(q[a + 3:a + 7])
                continue
            Rp = np.eye(3) if b.parent < 0 else xmat[b.parent]
            pp = np.zeros(3) if b.parent < 0 else xpos[b.parent]
            R = Rp @ quat_to_mat(b.quat)
            p = pp + Rp @ np.asarray(b.pos, dtype=np.float64)
            if b.jtype == HINGE:
                R = R @ axis_angle_mat(b.axis, q[self.qadr[i]])
            elif b.jtype == SLIDE:
                p = p + R @ (np.asarray(b.axis, dtype=np.float64) * q[self.qadr[i]])
            xpos[i], xmat[i] = p, R
        return xpos, xmat

    def _ancestors(self, i):
        out = []
        while i >= 0:
            out.append(i)
            i = self.bodies[i].parent
        return out

    def jacobian(self, q, body, point_world):
        """Translational / rotational Jacobian (3,nv) of a world point attached to `body`."""
        xpos, xmat = self.fk(q)
        Jv = np.zeros((3, self.nv))
        Jw = np.zeros((3, self.nv))
        for j in self._ancestors(body):
            b = self.bodies[j]
            d = self.dofadr[j]
            if b.jtype == HINGE:
                k = xmat[j] @ np.asarray(b.axis, dtype=np.float64)
                Jw[:, d] = k
                Jv[:, d] = np.cross(k, point_world - xpos[j])
            elif b.jtype == SLIDE:
                Jv[:, d] = xmat[j] @ np.asarray(b.axis, dtype=np.float64)
            elif b.jtype == FREE:
                Jv[:, d:d + 3] = np.eye(3)
                for c in range(3):
                    k = xmat[j][:, c] if GENESIS_COMPAT_FREE_ANG_BODYFRAME else np.eye(3)[:, c]
                    Jw[:, d + 3 + c] = k
                    Jv[:, d + 3 + c] = np.cross(k, point_world - xpos[j])
        return Jv, Jw

    def mass_matrix(self, q, with_armature=True):
        xpos, xmat = self.fk(q)
        M = np.zeros((self.nv, self.nv))
        for i, b in enumerate(self.bodies):
            if b.mass <= 0:
                continue
            c = xpos[i] + xmat[i] @ np.asarray(b.com, dtype=np.float64)
            Jv, Jw = self.jacobian(q, i, c)
            Iw = xmat[i] @ inertia6_to_mat(b.inertia) @ xmat[i].T
            M += b.mass * Jv.T @ Jv + Jw.T @ Iw @ Jw
        if with_armature:
            M += np.diag(self.dof_table()["armature"])
        return M

    def invweight0(self):
        """MuJoCo-style inverse weights at qpos0: per body (translational) and per dof."""
        q0 = self.full_q(self.qpos0)
        Minv = np.linalg.inv(self.mass_matrix(q0))
        xpos, xmat = self.fk(q0)
        body_w = np.zeros(self.nb)
        for i, b in enumerate(self.bodies):
            if self.dofadr[i] < 0 and all(self.dofadr[j] < 0 for j in self._ancestors(i)):
                continue        # welded to the world
            c = xpos[i] + xmat[i] @ np.asarray(b.com, dtype=np.float64)
            Jv, _ = self.jacobian(q0, i, c)
            body_w[i] = max(np.trace(Jv @ Minv @ Jv.T) / 3.0, GENESIS_COMPAT_MINVAL)
        dof_w = np.diag(Minv).copy()
        return body_w, dof_w

    # ---- flat arrays for the generic CPU oracle ----------------------------------------------
    def flat(self):
        """All model data as flat float64/int32 arrays, values pre-rounded to float32 so that the
        fp64 oracle and the fp32 CUDA path start from bit-identical constants."""
        f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).astype(np.float64))
        B = self.bodies
        dt = self.dof_table()
        body_w, dof_w = self.invweight0()
        out = dict(
            nb=self.nb, nv=self.nv, nq=self.nq, nu=self.nu, ngeom=len(self.geoms),
            eef_body=self.eef_body, cube_body=self.cube_body,
            n_obj=self.n_obj, task={"pick": 0, "stack": 1}[self.task],
            obj_body=np.array(self.obj_bodies, dtype=np.int32),
            obj_spawn=f32(self.obj_spawn).ravel(),
            task_opt=f32([self.stack_xy, self.stack_z, self.stack_min_dist, float(self.agent_mode == "qpos")]),
            parent=np.array([b.parent for b in B], dtype=np.int32),
            jtype=np.array([b.jtype for b in B], dtype=np.int32),
            dofadr=self.dofadr.copy(), qadr=self.qadr.copy(),
            pos=f32([b.pos for b in B]).ravel(),
            quat=f32([quat_normalize(b.quat) for b in B]).ravel(),
            axis=f32([b.axis for b in B]).ravel(),
            mass=f32([b.mass for b in B]),
            com=f32([b.com for b in B]).ravel(),
            inertia=f32([b.inertia for b in B]).ravel(),
            body_invweight=f32(body_w), dof_invweight=f32(dof_w),
            armature=f32(dt["armature"]), damping=f32(dt["damping"]), kp=f32(dt["kp"]), kv=f32(dt["kv"]),
            frc_lo=f32(dt["frc_lo"]), frc_hi=f32(dt["frc_hi"]),
            lim_lo=f32(dt["lim_lo"]), lim_hi=f32(dt["lim_hi"]),
            limited=dt["limited"], ctrl=dt["ctrl"],
            geom_body=np.array([g.body for g in self.geoms], dtype=np.int32),
            geom_group=np.array([g.group for g in self.geoms], dtype=np.int32),
            geom_half=f32([g.half for g in self.geoms]).ravel(),
            geom_pos=f32([g.pos for g in self.geoms]).ravel(),
            geom_quat=f32([quat_normalize(g.quat) for g in self.geoms]).ravel(),
            opt=f32([self.dt, self.gravity, self.plane_z, self.friction,
                     GENESIS_COMPAT_TIMECONST, GENESIS_COMPAT_DAMPRATIO, GENESIS_COMPAT_DMIN,
                     GENESIS_COMPAT_DMAX, GENESIS_COMPAT_WIDTH, GENESIS_COMPAT_MID, GENESIS_COMPAT_POWER,
                     GENESIS_COMPAT_MINVAL, self.success_z, self.spawn_lo[0], self.spawn_hi[0],
                     self.spawn_lo[1], self.spawn_hi[1], self.spawn_z, self.reg_scale()]),
            spawn_quat=f32(self.spawn_quat), home_q=f32(self.home_q),
            max_contacts=self.max_contacts,
        )
        return out


# ----------------------------------------------------------------------------------------------
# Franka Emika Panda + cube + plane  (reference franka/cube_pick.py:50-54)
# ----------------------------------------------------------------------------------------------
_PANDA_KP = [4500.0, 4500.0, 3500.0, 3500.0, 2000.0, 2000.0, 2000.0, 100.0, 100.0]   # franka/cube_stack.py:156
_PANDA_KV = [450.0, 450.0, 350.0, 350.0, 200.0, 200.0, 200.0, 10.0, 10.0]            # franka/cube_stack.py:157
_PANDA_FRC = [87.0] * 7 + [100.0] * 2                                                 # franka/cube_stack.py:158-161
FRANKA_HOME = np.array([0.0, -0.4, 0.0, -2.2, 0.0, 2.0, 0.8, 0.04, 0.04])             # franka/cube_pick.py:100

# name, parent, pos, quat(wxyz), range, mass, com, inertia(xx yy zz xy xz yz)
_PANDA_LINKS = [
    ("link1", 0, (0, 0, 0.333), (1, 0, 0, 0), (-2.8973, 2.8973), 4.970684, (0.003875, 0.002081, -0.04762),
     (0.70337, 0.70661, 0.0091170, -0.00013900, 0.0067720, 0.019169)),
    ("link2", 1, (0, 0, 0), (1, -1, 0, 0), (-1.7628, 1.7628), 0.646926, (-0.003141, -0.02872, 0.003495),
     (0.0079620, 0.028110, 0.025995, -0.003925, 0.010254, 0.000704)),
    ("link3", 2, (0, -0.316, 0), (1, 1, 0, 0), (-2.8973, 2.8973), 3.228604, (0.027518, 0.039252, -0.066502),
     (0.037242, 0.036155, 0.01083, -0.004761, -0.011396, -0.012805)),
    ("link4", 3, (0.0825, 0, 0), (1, 1, 0, 0), (-3.0718, -0.0698), 3.587895, (-0.05317, 0.104419, 0.027454),
     (0.025853, 0.019552, 0.028323, 0.007796, -0.001332, 0.008641)),
    ("link5", 4, (-0.0825, 0.384, 0), (1, -1, 0, 0), (-2.8973, 2.8973), 1.225946, (-0.011953, 0.041065, -0.038437),
     (0.035549, 0.029474, 0.008627, -0.002117, -0.004037, 0.000229)),
    ("link6", 5, (0, 0, 0), (1, 1, 0, 0), (-0.0175, 3.7525), 1.666555, (0.060149, -0.014117, -0.010517),
     (0.001964, 0.004354, 0.005433, 0.000109, -0.001158, 0.000341)),
    ("link7", 6, (0.088, 0, 0), (1, 1, 0, 0), (-2.8973, 2.8973), 0.735522, (0.010517, -0.004252, 0.061597),
     (0.012516, 0.010027, 0.004815, -0.000428, -0.001196, -0.000741)),
]

# five pad boxes per finger: (half sizes, finger-local centre)  -- SURVEY.md Appendix B
_PANDA_PADS = [
    ((0.0085, 0.004, 0.0085), (0.0, 0.0055, 0.0445)),
    ((0.003, 0.002, 0.003), (0.0055, 0.002, 0.05)),
    ((0.003, 0.002, 0.003), (-0.0055, 0.002, 0.05)),
    ((0.003, 0.002, 0.0035), (0.0055, 0.002, 0.0395)),
    ((0.003, 0.002, 0.0035), (-0.0055, 0.002, 0.0395)),
]

CUBE_SIZE = 0.04                                          # franka/cube_pick.py:53


def _cube_body(size=CUBE_SIZE, density=GENESIS_COMPAT_RIGID_DENSITY):
    m = density * size ** 3
    i = m * size * size / 6.0
    return Body("cube", -1, (0.65, 0.0, size / 2), (1, 0, 0, 0), jtype=FREE, mass=m, inertia=(i, i, i, 0, 0, 0))


def franka_bodies(scale: float = 1.0, base_pos=(0.0, 0.0, 0.0), wrist_force: float = 87.0):
    """Panda body table (link0 .. link7, hand, two fingers) and its pad geoms.  `scale` is the MJCF morph scale of the kitchen
    scenes (`tasks/utils.py:369-376`: scale 0.6 at (-0.5, 0, 0.7)): lengths x s, masses x s^3, inertias x s^5; joint ranges, armature,
    damping and the task's explicit gains are kept ([UPSTREAM-RECALL]: Genesis scales geometry and inertial data of a morph)."""
    s = float(scale)
    sc = lambda v: tuple(float(x) * s for x in v)
    bodies = [Body("link0", -1, tuple(float(x) for x in base_pos), (1, 0, 0, 0), mass=0.629769 * s ** 3, com=sc((-0.041018, -0.00014, 0.049974)),
                   inertia=tuple(i * s ** 5 for i in (0.00315, 0.00388, 0.004285, 8.2904e-7, 0.00015, 8.2299e-6)))]
    for j, (name, parent, pos, quat, rng, mass, com, inertia) in enumerate(_PANDA_LINKS):
        bodies.append(Body(name, parent, sc(pos), quat, jtype=HINGE, axis=(0, 0, 1), range=rng, limited=True,
                           armature=0.1, damping=1.0, kp=_PANDA_KP[j], kv=_PANDA_KV[j],
                           frc=((-wrist_force, wrist_force) if j >= 4 else (-_PANDA_FRC[j], _PANDA_FRC[j])), mass=mass * s ** 3, com=sc(com),
                           inertia=tuple(i * s ** 5 for i in inertia)))
    bodies.append(Body("hand", 7, sc((0, 0, 0.107)), (0.9238795, 0, 0, -0.3826834), mass=0.73 * s ** 3, com=sc((-0.01, 0, 0.03)),
                       inertia=tuple(i * s ** 5 for i in (0.001, 0.0025, 0.0017, 0, 0, 0))))
    for f, (name, quat) in enumerate((("left_finger", (1, 0, 0, 0)), ("right_finger", (0, 0, 0, 1)))):
        bodies.append(Body(name, 8, sc((0, 0, 0.0584)), quat, jtype=SLIDE, axis=(0, 1, 0), range=(0.0, 0.04 * s), limited=True,
                           armature=0.1, damping=1.0, kp=_PANDA_KP[7 + f], kv=_PANDA_KV[7 + f],
                           frc=(-_PANDA_FRC[7 + f], _PANDA_FRC[7 + f]), mass=0.015 * s ** 3,
                           inertia=tuple(i * s ** 5 for i in (2.375e-6, 2.375e-6, 7.5e-7, 0, 0, 0))))
    geoms = []
    for fb in (9, 10):
        for half, pos in _PANDA_PADS:
            geoms.append(BoxGeom(fb, sc(half), sc(pos), group=0))
    return bodies, geoms


def franka_cube_pick_model(wrist_force: float = None, pyramid_mu_scaling: bool = None) -> SceneModel:
    """Plane + Panda (9 dofs) + free 4 cm cube: the scene of `franka/cube_pick.py:50-54`.
    wrist_force / pyramid_mu_scaling: the two switchable Genesis-compatibility risks (GENESIS_COMPAT_PANDA_WRIST_FORCE,
    GENESIS_COMPAT_PYRAMID_MU_SCALING; None = the module constants)."""
    bodies, geoms = franka_bodies(wrist_force=GENESIS_COMPAT_PANDA_WRIST_FORCE if wrist_force is None else float(wrist_force))
    bodies.append(_cube_body())
    geoms.append(BoxGeom(11, (CUBE_SIZE / 2,) * 3, group=1))
    return SceneModel("franka", bodies, geoms, eef_body=8, cube_body=11, home_q=FRANKA_HOME.copy(), n_arm=7, n_fing=2,
                      pyramid_mu_scaling=GENESIS_COMPAT_PYRAMID_MU_SCALING if pyramid_mu_scaling is None else bool(pyramid_mu_scaling))


# cube_1 (picked), cube_2 (target), three distractors: body positions of franka/cube_stack.py:54-77 (the distractors'
# build-time positions are unseeded np.random draws there, overwritten by every reset; fixed representative values here)
_STACK_CUBE_POS = [(0.6, -0.1, 0.02), (0.45, 0.15, 0.02), (0.35, -0.25, 0.02), (0.55, 0.25, 0.02), (0.65, 0.05, 0.02)]
# spawn ranges (x_lo, x_hi, y_lo, y_hi): cube_1 franka/cube_stack.py:125-126, cube_2 :132-133, distractors :141-142
_STACK_SPAWN = [(0.45, 0.75, -0.2, 0.2), (0.3, 0.7, -0.3, 0.3), (0.3, 0.7, -0.3, 0.3), (0.3, 0.7, -0.3, 0.3), (0.3, 0.7, -0.3, 0.3)]


def franka_cube_stack_model(n_distractors: int = 3) -> SceneModel:
    """Plane + Panda + cube_1 + cube_2 + 3 distractor cubes, all free 4 cm boxes: `franka/cube_stack.py:42-96`
    (`FrankaCubeStackEmptyBatch`).  Gains are the explicit ones of `:156-161` (already the Panda table here)."""
    base = franka_cube_pick_model()
    bodies = base.bodies[:-1]
    geoms = [g for g in base.geoms if g.group == 0]
    n_obj = 2 + n_distractors
    obj = []
    for k in range(n_obj):
        b = _cube_body()
        b.name = ("cube_1", "cube_2")[k] if k < 2 else f"distractor_{k - 2}"
        b.pos = _STACK_CUBE_POS[k]
        obj.append(len(bodies))
        bodies.append(b)
        geoms.append(BoxGeom(obj[-1], (CUBE_SIZE / 2,) * 3, group=1))
    return SceneModel("franka_stack", bodies, geoms, eef_body=8, cube_body=obj[0], home_q=FRANKA_HOME.copy(), n_arm=7, n_fing=2,
                      obj_bodies=obj, obj_spawn=_STACK_SPAWN[:n_obj], task="stack",
                      max_contacts=GENESIS_COMPAT_MAX_CONTACTS_STACK)


KITCHEN_ISLAND_TOP_Z = 0.7000312834978104     # tasks/utils.py:378-384 (= -0.2003899186849594 + GLB bounding-box top)
# cube_1, cube_2 (franka/cube_stack_kitchen_batch.py:71-72, 78-79), distractors (:87-88): x_lo, x_hi, y_lo, y_hi
_KITCHEN_SPAWN = [(-0.3, -0.1, -0.15, 0.15), (-0.3, -0.1, -0.15, 0.15), (-0.35, 0.0, -0.2, 0.2), (-0.35, 0.0, -0.2, 0.2), (-0.35, 0.0, -0.2, 0.2)]
_KITCHEN_POS = [(0.1, 0.0), (-0.1, 0.05), (-0.2, -0.15), (0.15, 0.2), (-0.25, 0.1)]       # utils.py:386-425 (distractors: unseeded draws)


def franka_cube_stack_kitchen_model() -> SceneModel:
    """`FrankaCubeStackKitchenBatch` (`franka/cube_stack_kitchen_batch.py:26-150`, scene `tasks/utils.py:239-427`): what the
    reference wires to `CubeStack-v0` + franka (`env.py:115`).  Panda at (-0.5, 0, 0.7) with morph scale 0.6, five free 4 cm cubes
    on the kitchen island (its top modelled as a plane, as for the SO-101 scenes; the rest of the kitchen has `collision=False`),
    same gains, observation and reward as the plane variant."""
    bodies, geoms = franka_bodies(0.6, (-0.5, 0.0, 0.7))
    z = KITCHEN_ISLAND_TOP_Z + 0.02 + 0.001                                      # cube_stack_kitchen_batch.py:67
    obj = []
    for k in range(5):
        b = _cube_body()
        b.name = ("cube_1", "cube_2")[k] if k < 2 else f"distractor_{k - 2}"
        b.pos = (*_KITCHEN_POS[k], z)
        obj.append(len(bodies))
        bodies.append(b)
        geoms.append(BoxGeom(obj[-1], (CUBE_SIZE / 2,) * 3, group=1))
    # reset writes the literal finger opening 0.04 (:94), beyond the scaled finger range (0.024): the joint-limit rows push it
    # back within a few steps, as they would upstream
    return SceneModel("franka_stack_kitchen", bodies, geoms, eef_body=8, cube_body=obj[0], home_q=FRANKA_HOME.copy(), n_arm=7, n_fing=2,
                      plane_z=KITCHEN_ISLAND_TOP_Z, spawn_z=z, obj_bodies=obj, obj_spawn=_KITCHEN_SPAWN, task="stack",
                      max_contacts=GENESIS_COMPAT_MAX_CONTACTS_STACK)


def get_model(robot: str, task: str = "cube_pick", scene: str = "kitchen") -> SceneModel:
    if robot == "franka" and task == "cube_stack":
        return franka_cube_stack_kitchen_model() if scene == "kitchen" else franka_cube_stack_model()
    if robot == "franka":
        return franka_cube_pick_model()
    if robot == "so101" and task == "cube_stack":
        from .model_so101 import so101_cube_stack_model
        return so101_cube_stack_model()
    if robot == "so101":
        from .model_so101 import so101_cube_pick_model
        return so101_cube_pick_model()
    raise NotImplementedError(robot)
