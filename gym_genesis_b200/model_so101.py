"""SO-101 arm + cube + supporting plane (BASELINE.json config 4).

Constants restate `assets/robots/SO-ARM100/Simulation/SO101/so101_old_calib.xml` of the reference (line numbers
below refer to that file); the task overrides come from `gym_genesis/tasks/so101/cube_pick.py:38-42` (friction 5 on
robot and cube, arm kp 1000 / kv 200, jaw keeps the MJCF kp 17.8 / kv 0) and the joint class `sts3215`
(`:22-23`: damping 0.60, armature 0.028, forcerange +-3.35 `:152-157`).  `frictionloss` (0.052) is not modelled.

Placement follows SURVEY.md section 8d / Appendix C: the reference's *pick* placement (scale 4.0, no yaw,
`tasks/utils.py:559-566`) puts the gripper 1.45 m away from the cube, so the batched config uses the *stack* placement
(`tasks/utils.py:730-744`: base (-0.5, 0, 0.7), yaw 90 deg, scale 1.3, home deg2rad(0,-177,165,72,-83,0)
`so101/cube_stack_batch.py:106`) with the pick task's spawn, obs and reward; `placement="pick_literal"` keeps the literal
one.  The island top is modelled as a plane at z = 0.7000312834978104 (`tasks/utils.py:571-578`).

Collision: the 13 collision geoms are STL meshes (17k-54k triangles) that Genesis convexifies; here only the two jaws
collide and each is approximated by ONE box fitted to the finger part of its mesh (moving_jaw_so101_v1.stl,
wrist_roll_follower_so101_v1.stl) -- an explicit, documented primitive approximation.
"""
from __future__ import annotations

import numpy as np

from .model import (BoxGeom, Body, FIXED, HINGE, SceneModel, _cube_body, CUBE_SIZE, quat_mul, quat_normalize)

ISLAND_TOP_Z = 0.7000312834978104           # tasks/utils.py:571-578 (= -0.2003899186849594 + GLB bounding-box top)
SO101_STACK_HOME = np.deg2rad([0.0, -177.0, 165.0, 72.0, -83.0, 0.0])     # so101/cube_stack_batch.py:106

# name, parent, pos, quat(wxyz), range, mass, com, fullinertia (xx yy zz xy xz yz)            xml line
_SO101_BODIES = [
    ("base", -1, (0, 0, 0), (1, 0, 0, 0), None, 0.147, (0.020739, 0.00204287, 0.065966),
     (0.000136117, 0.000114686, 0.000130364, 4.59787e-07, 9.75275e-08, -4.97151e-06)),                       # :32-33
    ("shoulder", 0, (0.0207909, -0.0230745, 0.0948817), (7.88629e-16, -0.707107, -0.707107, 7.69003e-16),
     (-1.9198621771937616, 1.9198621771937634), 0.100006, (-0.0307604, -1.66727e-05, -0.0252713),
     (8.3759e-05, 8.10403e-05, 2.39783e-05, 7.55525e-08, -1.16342e-06, 1.54663e-07)),                        # :45-48
    ("upper_arm", 1, (-0.0303992, -0.0182778, -0.0542), (0.5, -0.5, 0.5, 0.5),
     (-3.31612557878923, 0.1745329251994294), 0.103, (-0.00838224, 0.0898471, 0.0184089),
     (0.000147318, 4.08002e-05, 0.000142487, 1.97819e-05, 8.97326e-09, 4.03016e-08)),                        # :59-62
    ("lower_arm", 2, (-0.028, 0.11257, 1.24466e-16), (2.34133e-16, -3.33458e-16, -3.62568e-16, -1.0),
     (-0.17453292519941607, 3.14159265358981), 0.104, (-0.00324376, -0.0980701, 0.0182831),
     (0.000159844, 2.87438e-05, 0.00014529, -7.41152e-06, 4.90188e-08, 1.26409e-06)),                        # :70-73
    ("wrist", 3, (-0.0052, -0.1349, 3.40439e-16), (1.0, -3.94784e-16, 1.66357e-15, 3.78078e-09),
     (-1.6580627969561947, 1.6580627818330316), 0.079, (-0.000103312, -0.0386143, 0.0281156),
     (3.68263e-05, 2.5391e-05, 2.1e-05, 1.7893e-08, -5.28128e-08, 3.6412e-06)),                              # :84-87
    ("gripper", 4, (2.77556e-17, -0.0611, 0.0181), (3.31663e-08, -3.31663e-08, -0.707107, -0.707107),
     (-2.7925268969992394, 2.792526709382615), 0.087, (0.000213627, 0.000245138, -0.025187),
     (2.75087e-05, 4.33657e-05, 3.45059e-05, -3.35241e-07, -5.7352e-06, -5.17847e-08)),                      # :95-98
    ("moving_jaw_so101_v1", 5, (0.0202, 0.0188, -0.0234), (0.707107, 0.707107, 0.0, -4.88629e-17),
     (-0.17453292519941963, 1.7453292519943429), 0.012, (-0.00157495, -0.0300244, 0.0192755),
     (6.61427e-06, 1.89032e-06, 5.28738e-06, -3.19807e-07, -5.90717e-09, -1.09945e-07)),                     # :108-111
]
_DAMPING, _ARMATURE, _FRC = 0.60, 0.028, 3.35           # :22, :152-157
_ARM_KP, _ARM_KV, _JAW_KP, _JAW_KV = 1000.0, 200.0, 17.8, 0.0   # so101/cube_pick.py:41-42 ; xml :23, joints_properties kv = 0

# one box per jaw, fitted to the finger part of the collision mesh, expressed in the BODY frame at scale 1
_FIXED_JAW_BOX = ((0.0105, 0.008, 0.0327), (-0.0195, -0.000218214, -0.07175))      # on body `gripper` (geom :104)
_MOVING_JAW_BOX = ((0.0075, 0.0285, 0.0075), (-0.003, -0.0535, 0.0189))            # on body `moving_jaw_so101_v1` (geom :114)


def so101_bodies(scale=1.0, base_pos=(0.0, 0.0, 0.0), base_yaw_deg=0.0, friction_gains=True, arm_gains=(_ARM_KP, _ARM_KV)):
    s = float(scale)
    yaw = np.deg2rad(base_yaw_deg)
    base_quat = (np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2))
    bodies = []
    for k, (name, parent, pos, quat, rng, mass, com, inertia) in enumerate(_SO101_BODIES):
        kw = dict(mass=mass * s ** 3, com=tuple(c * s for c in com), inertia=tuple(i * s ** 5 for i in inertia))
        if k == 0:
            bodies.append(Body(name, -1, tuple(base_pos), quat_mul(base_quat, quat_normalize(quat)), jtype=FIXED, **kw))
            continue
        arm = k <= 5
        bodies.append(Body(name, parent, tuple(p * s for p in pos), quat, jtype=HINGE, axis=(0, 0, 1), range=rng, limited=True,
                           armature=_ARMATURE, damping=_DAMPING, kp=arm_gains[0] if arm else _JAW_KP, kv=arm_gains[1] if arm else _JAW_KV,
                           frc=(-_FRC, _FRC), **kw))
    return bodies


def so101_cube_pick_model(placement="stack", success_height=None, pyramid_mu_scaling=None) -> SceneModel:
    if placement == "stack":
        scale, base_pos, yaw, home = 1.3, (-0.5, 0.0, 0.7), 90.0, SO101_STACK_HOME.copy()
    elif placement == "pick_literal":
        scale, base_pos, yaw, home = 4.0, (-0.5, 0.0, 0.7), 0.0, np.zeros(6)        # tasks/utils.py:559-566, so101/cube_pick.py:74
    else:
        raise ValueError(placement)
    bodies = so101_bodies(scale, base_pos, yaw)
    cube = _cube_body()
    cube.pos = (-0.3, 0.0, ISLAND_TOP_Z + CUBE_SIZE / 2)
    bodies.append(cube)
    geoms = [BoxGeom(5, tuple(h * scale for h in _FIXED_JAW_BOX[0]), tuple(p * scale for p in _FIXED_JAW_BOX[1]), group=0),
             BoxGeom(6, tuple(h * scale for h in _MOVING_JAW_BOX[0]), tuple(p * scale for p in _MOVING_JAW_BOX[1]), group=0),
             BoxGeom(7, (CUBE_SIZE / 2,) * 3, group=1)]
    sm = SceneModel("so101", bodies, geoms, eef_body=5, cube_body=7, home_q=home, n_arm=5, n_fing=1,
                    plane_z=ISLAND_TOP_Z, friction=5.0,                                    # so101/cube_pick.py:38-39
                    spawn_lo=(-0.32, -0.05), spawn_hi=(-0.28, 0.05),                       # so101/cube_pick.py:61-62
                    spawn_z=ISLAND_TOP_Z + 0.02 + 0.001, spawn_quat=(1.0, 0.0, 0.0, 0.0),  # so101/cube_pick.py:66,68
                    success_z=0.1 if success_height is None else float(success_height),    # so101/cube_pick.py:112 (literal; see D11)
                    settle_steps=0)                                                        # so101/cube_pick.py:81 (commented out)
    if pyramid_mu_scaling is not None:                     # GENESIS_COMPAT_PYRAMID_MU_SCALING override (model.py)
        sm.pyramid_mu_scaling = bool(pyramid_mu_scaling)
    return sm


# so101/cube_stack_batch.py:72-75 (cube_1, cube_2) and :93-94 (distractors): x_lo, x_hi, y_lo, y_hi
_SO101_STACK_SPAWN = [(-0.3, -0.1, -0.1, 0.1), (-0.3, -0.1, -0.1, 0.1), (-0.35, 0.0, -0.2, 0.2), (-0.35, 0.0, -0.2, 0.2), (-0.35, 0.0, -0.2, 0.2)]
# build-time positions, overwritten by every reset (tasks/utils.py:756-790: cube_1, cube_2 literal; distractors unseeded draws)
_SO101_STACK_POS = [(0.1, 0.0), (-0.1, 0.05), (-0.2, -0.15), (0.15, 0.2), (-0.3, 0.15)]


def so101_cube_stack_model() -> SceneModel:
    """SO-101 `CubeStackBatch` (`so101/cube_stack_batch.py:25-226`, scene `tasks/utils.py:593-794`): robot at (-0.5, 0, 0.7),
    yaw 90 deg, scale 1.3 (`utils.py:733-744`), five free 4 cm cubes on the island top (modelled as a plane at ISLAND_TOP_Z),
    MJCF actuator gains (kp 17.8, kv 0: the task's set_dofs_kp / set_friction calls are commented out, `:46-49`), default
    friction, one settle step in reset (`:111`), agent_pos = joint positions (`:176`)."""
    scale = 1.3
    bodies = so101_bodies(scale, (-0.5, 0.0, 0.7), 90.0, arm_gains=(_JAW_KP, _JAW_KV))
    geoms = [BoxGeom(5, tuple(h * scale for h in _FIXED_JAW_BOX[0]), tuple(p * scale for p in _FIXED_JAW_BOX[1]), group=0),
             BoxGeom(6, tuple(h * scale for h in _MOVING_JAW_BOX[0]), tuple(p * scale for p in _MOVING_JAW_BOX[1]), group=0)]
    z = ISLAND_TOP_Z + 0.02 + 0.001                                               # utils.py:752, cube_stack_batch.py:67
    obj = []
    for k in range(5):
        b = _cube_body()
        b.name = ("cube_1", "cube_2")[k] if k < 2 else f"distractor_{k - 2}"
        b.pos = (*_SO101_STACK_POS[k], z)
        obj.append(len(bodies))
        bodies.append(b)
        geoms.append(BoxGeom(obj[-1], (CUBE_SIZE / 2,) * 3, group=1))
    from .model import GENESIS_COMPAT_MAX_CONTACTS_STACK
    return SceneModel("so101_stack", bodies, geoms, eef_body=5, cube_body=obj[0], home_q=SO101_STACK_HOME.copy(), n_arm=5, n_fing=1,
                      plane_z=ISLAND_TOP_Z, spawn_z=z, spawn_quat=(0.0, 0.0, 0.0, 1.0),   # cube_stack_batch.py:66
                      settle_steps=1, obj_bodies=obj, obj_spawn=_SO101_STACK_SPAWN, task="stack", stack_min_dist=0.06,
                      agent_mode="qpos", max_contacts=GENESIS_COMPAT_MAX_CONTACTS_STACK)
