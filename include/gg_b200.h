/* gg_b200.h -- C ABI of the B200-native batched env.step for the gym-genesis manipulation tasks.
 *
 * The reference has no FFI: its hot path is Python calling the Genesis entity API.  Each entry
 * point below replaces one group of those calls (citations are into /root/reference):
 *
 *   gg_create        <- scene.build(n_envs=N)                      gym_genesis/tasks/franka/cube_pick.py:65
 *                       (+ model upload: add_entity Plane/MJCF/Box  :50-54, SimOptions dt=0.01 :45)
 *   gg_bind_state    <- Genesis-internal field allocation (same call; torch owns the memory here)
 *   gg_reset         <- cube.set_pos/set_quat, franka.set_qpos(zero_velocity=True),
 *                       control_dofs_position(home) x2, scene.step(), get_obs()          :86-112
 *   gg_step          <- control_dofs_position(action[:, :7]) / (action[:, 7:]), scene.step(),
 *                       compute_reward(), get_obs()                                     :122-157
 *                       and GenesisEnv.step's is_success / terminated / truncated        gym_genesis/env.py:61-69
 *   gg_get_state /
 *   gg_set_state     <- entity.get_qpos / get_dofs_velocity / set_qpos / set_dofs_velocity (state injection,
 *                       used by the parity harness)
 *   gg_step_host     <- the same calls with host numpy arrays as the reference passes / receives them
 *                       (README.md:34 numpy action; cube_pick.py:130-135 numpy reward; env.py:63-65 numpy flags)
 *   gg_stats         <- (new) per-shard episode statistics for the optional NCCL all-gather
 *   gg_ik            <- robot.inverse_kinematics(link=eef, pos=, quat=, envs_idx=)   examples/franka/pick_cube_state.py:46-51,
 *                       examples/franka/stack_cube_state.py:78-83, examples/so_101/collect_task_stack_cube_batch.py:88-94
 *                       (the caller-side step of every expert policy)
 *   gg_record /
 *   gg_export_successful <- the collectors' per-step `.detach().cpu().numpy()` appends and the "save envs whose reward was ever
 *                       > 0" loop                                          examples/franka/pick_cube_state.py:86-113
 *                       (frames stay on the device; one compacted copy per episode)
 *   gg_stack_*      <- the same calls of the five-cube stacking task:
 *                       gym_genesis/tasks/franka/cube_stack.py:42-96 (scene: cube_1, cube_2, 3 distractors, Panda),
 *                       :119-170 (reset: per-cube spawn ranges, home pose, explicit gains, one settle step),
 *                       :180-186 (step), :188-196 (reward xy_dist < 0.05 & z_diff > 0.03), :198-216 (obs 9 + 14)
 *
 * Conventions: plain pointers and sizes only; every pointer argument of gg_step / gg_reset is a
 * DEVICE pointer into caller-owned (torch) memory unless stated otherwise; all work is enqueued
 * asynchronously on `stream`; no allocation and no host synchronisation after gg_create.
 * Return value: 0 = ok, negative = error (message via gg_last_error()).
 */
#ifndef GG_B200_H
#define GG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_MAX_ARM 7        /* serial hinge chain (Panda 7, SO-101 5)                     */
#define GG_MAX_FING 2       /* gripper dofs hanging off the last arm link (Panda 2 slides) */
#define GG_MAX_DOF 9
#define GG_MAX_PADS 12      /* box collision geoms on the gripper                          */
#define GG_MAX_CONTACTS 64  /* per-env contact cap (deterministic order)    */
#define GG_ENV_STATE_DIM 11 /* cube pos3 + quat4 + (eef - cube)3 + |.|1                    */

typedef struct CUstream_st* gg_stream_t;
typedef struct GgSim GgSim;

/* Compiled scene constants (host fills this once; copied to the device at gg_create).
 * Arm link l is a hinge about its own local z axis; the last arm link carries its fixed children
 * (Panda "hand") merged into mass/com/inertia.  Fingers are children of the last arm link. */
typedef struct GgModel {
    int32_t n_arm, n_fing, fing_hinge, n_pads;
    int32_t settle_steps, horizon, max_contacts, reserved0;
    float base_pos[3], base_quat[4];
    float link_pos[GG_MAX_ARM][3], link_quat[GG_MAX_ARM][4];
    float link_mass[GG_MAX_ARM], link_com[GG_MAX_ARM][3], link_inertia[GG_MAX_ARM][6]; /* xx yy zz xy xz yz */
    float fing_pos[GG_MAX_FING][3], fing_quat[GG_MAX_FING][4], fing_axis[GG_MAX_FING][3];
    float fing_mass[GG_MAX_FING], fing_com[GG_MAX_FING][3], fing_inertia[GG_MAX_FING][6];
    float eef_pos[3], eef_quat[4];               /* observed end-effector frame in the last arm link */
    float armature[GG_MAX_DOF], damping[GG_MAX_DOF], kp[GG_MAX_DOF], kv[GG_MAX_DOF];
    float frc_lo[GG_MAX_DOF], frc_hi[GG_MAX_DOF], lim_lo[GG_MAX_DOF], lim_hi[GG_MAX_DOF];
    float dof_invweight[GG_MAX_DOF];
    float body_invweight[1 + GG_MAX_FING];        /* last arm link, finger 0, finger 1 */
    int32_t pad_body[GG_MAX_PADS];                /* 0 = last arm link, 1 + f = finger f */
    float pad_half[GG_MAX_PADS][3], pad_pos[GG_MAX_PADS][3], pad_quat[GG_MAX_PADS][4];
    float pad_bound[1 + GG_MAX_FING][4];          /* per pad-carrying body: bounding sphere (centre xyz, radius; r < 0: no pads) */
    float cube_half[3], cube_mass, cube_inertia[3], cube_invweight;
    float dt, gravity, plane_z, friction;
    float tc, dr, dmin, dmax, width, mid, power, minval;   /* solref / solimp */
    float home_q[GG_MAX_DOF];
    float spawn_lo[2], spawn_hi[2], spawn_z, spawn_quat[4], success_z;
    float reg_scale;   /* regulariser of a pyramidal contact row: R = reg_scale (1 - imp) / imp (w_a + w_b); the MuJoCo-lineage value is
                          2 mu^2 (1 + mu^2) (GENESIS_COMPAT_PYRAMID_MU_SCALING on), 1 switches the friction scaling off */
    float reserved1[3];
} GgModel;

/* Run-time options; defaults (all zero except iterations/tolerance) reproduce the reference. */
typedef struct GgConfig {
    int32_t solver_iterations;   /* Newton iteration cap per substep (default 10)                  */
    float solver_tolerance;      /* stop when max_i |g_i|/H_ii <= tol * (1 + |a|_inf) (default 1e-5)  */
    int32_t auto_reset;          /* 1: envs that finished (terminated|truncated) reset on next step */
    int32_t time_limit;          /* 1: truncated = step_count >= model.horizon (reference: always 0)*/
    float reset_prob;            /* per-env Bernoulli reset probability per step (benchmark C3)      */
    uint64_t seed;               /* Philox key for device-side spawn draws and random resets         */
    int64_t first_env_id;        /* global id of local env 0 (multi-GPU sharding)                    */
} GgConfig;

const char* gg_last_error(void);
int gg_version(void);

int gg_create(const GgModel* model, int n_envs, int device, GgSim** out);
int gg_destroy(GgSim* sim);
int gg_configure(GgSim* sim, const GgConfig* cfg);

/* State lives in ONE caller-owned float32 device buffer of gg_state_rows() x gg_state_stride()
 * (struct-of-arrays, env index fastest) plus two int32[n_envs] counters. */
int gg_state_rows(const GgSim* sim);
int gg_state_stride(const GgSim* sim);
int gg_bind_state(GgSim* sim, float* state, int32_t* step_count, int32_t* episode);

/* One env.step for every env.  action: (n_envs, n_dof) row-major absolute joint targets.
 * obs_agent: (n_envs, 7 + n_fing); obs_env: (n_envs, 11); reward: (n_envs); terminated /
 * truncated: (n_envs) bytes.  reset_mask (nullable): envs with a non-zero byte are reset (spawn
 * draw + home pose + settle step) instead of stepped. */
int gg_step(GgSim* sim, const float* action, float* obs_agent, float* obs_env, float* reward,
            uint8_t* terminated, uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream);

/* Host-resident I/O for callers whose policy lives on the CPU -- the reference's own calling convention: numpy actions in
 * (README.md:34), numpy reward / terminated / truncated out (cube_pick.py:130-135, env.py:63-65).  ZERO-COPY: `action` is
 * (n_envs, n_dof) in page-locked host memory and is read by the step kernel itself (cp.async over PCIe / C2C, overlapped
 * with the bias-force computation); the nullable mirrors are page-locked host arrays the kernel writes next to the device
 * outputs.  No copy-engine transfer and no second launch is involved.  wait != 0: the call returns after the stream has
 * drained, i.e. when the mirrors are readable (the only entry point that may synchronise). */
typedef struct GgHostIO {
    const float* action;
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    int32_t wait;
    int32_t reserved0;
    /* optional (NULL = off), gg_step_host only: page-locked host mirrors of the observations -- what a host-resident policy reads
     * every step (README.md:34 feeds `observation` to policy()) -- written by the same kernel with 128-bit stores ... */
    float* obs_agent;          /* (n_envs, 7 + n_fing) */
    float* obs_env;            /* (n_envs, 11)         */
    /* ... and a DEVICE buffer (n_envs, n_dof) that receives the action batch the step consumed: control_dofs_position targets
     * persist across steps (cube_pick.py:123-124), so `push()` / masked resets need them on the device after a host-fed step. */
    float* action_dev;
} GgHostIO;
int gg_step_host(GgSim* sim, const GgHostIO* io, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated,
                 uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream);

/* Reset envs selected by mask (NULL = all).  cube_xy (nullable): (n_envs, 2) explicit spawn
 * positions (the host RandomState stream of the reference); NULL -> device Philox draw keyed by
 * (cfg.seed, global env id, episode).  Runs model.settle_steps substeps holding the home target
 * for the reset envs, then writes observations for all envs. */
int gg_reset(GgSim* sim, const uint8_t* mask, const float* cube_xy, float* obs_agent, float* obs_env,
             gg_stream_t stream);

/* Env-major copies for state injection / inspection: q (n_envs, n_dof + 7), qd and qacc (n_envs, n_dof + 6). */
int gg_get_state(GgSim* sim, float* q, float* qd, float* qacc, gg_stream_t stream);
int gg_set_state(GgSim* sim, const float* q, const float* qd, const float* qacc, gg_stream_t stream);

/* Recompute observations / reward / terminated from the current state without stepping. */
int gg_observe(GgSim* sim, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated, gg_stream_t stream);

/* Device-side episode statistics accumulated by gg_step (8 doubles written to `out8`, DEVICE pointer):
 * [0] env-steps, [1] episodes finished, [2] successes, [3] return sum, [4] length sum,
 * [5] non-finite envs auto-reset, [6] solver iterations sum, [7] contacts sum. */
int gg_stats(GgSim* sim, double* out8, int clear, gg_stream_t stream);

/* Batched inverse kinematics of the observed end-effector link (damped least squares, Genesis'
 * `RigidEntity.inverse_kinematics` defaults below; random re-sampling of unsolved rows when max_samples > 1, see csrc/gg_ik.cuh).
 * Zero-initialise and call gg_ik_defaults() to get: max_iters 20, damping 0.01, pos_tol 5e-4, rot_tol 5e-3,
 * max_step 0.5, respect_limits 1, masks all 1. */
typedef struct GgIkConfig {
    int32_t max_iters;        /* max_solver_iters: DLS updates per env                                  */
    float damping;            /* J^T (J J^T + damping^2 I)^-1                                            */
    float pos_tol, rot_tol;   /* solved when |e_pos| < pos_tol and |e_rot| < rot_tol (m, rad)            */
    float max_step;           /* max_step_size: per-dof clamp of one update (rad)                        */
    int32_t respect_limits;   /* respect_joint_limit: clip to [lim_lo, lim_hi] after every update        */
    float pos_mask[3], rot_mask[3]; /* pos_mask / rot_mask: 1 = constrain this world axis, 0 = free      */
    int32_t max_samples;      /* max_samples: total attempts per row.  Attempt 0 starts from init_q / the current joints; attempt s >= 1
                                 of a row that is still unsolved starts from arm joints drawn uniformly inside the joint limits by the
                                 device Philox stream (key = seed, counter = (row, s, 'IKRS', dof / 4)).  The result is the FIRST attempt
                                 that meets the tolerances, else the attempt with the smallest max(|e_pos| / pos_tol, |e_rot| / rot_tol).
                                 gg_ik_defaults: 1 (Genesis' default is 50)                                  */
    uint32_t seed_lo, seed_hi;
} GgIkConfig;
void gg_ik_defaults(GgIkConfig* cfg);

/* n_rows IK problems.  Row r belongs to env envs_idx[r] (NULL: env r; then n_rows <= n_envs).
 * target_pos (n_rows, 3) / target_quat (n_rows, 4, wxyz): DEVICE pointers; either may be NULL (that half of the pose is
 * left free).  init_q (n_rows, n_dof) nullable: start from the env's current joint positions (Genesis init_qpos=None).
 * q_out (n_rows, n_dof): solution, gripper dofs copied from the start configuration.  err_out (n_rows, 2) nullable:
 * final |e_pos|, |e_rot|.  iters_out (n_rows) nullable: DLS updates applied.  The simulation state is not modified. */
int gg_ik(GgSim* sim, int n_rows, const int32_t* envs_idx, const float* target_pos, const float* target_quat,
          const float* init_q, float* q_out, float* err_out, int32_t* iters_out, const GgIkConfig* cfg, gg_stream_t stream);

/* Device-side trajectory store for the data collectors (examples/franka/pick_cube_state.py:86-113).
 * traj: caller-owned DEVICE buffer of (horizon, n_envs, F) float32 frames, F = gg_frame_dim() =
 * (7 + n_fing) + 11 + n_dof + 1, frame = [agent_pos | environment_state | action | reward].
 * ever: caller-owned DEVICE int32[n_envs]; gg_record sets ever[e] = 1 when env e's reward is > 0 (the collector's
 * `np.any(rewards[:, b] > 0)`); the caller zeroes it at episode start.
 * gg_record(t): frame t of every env <- the given observation / action / reward arrays (the outputs of the gg_step that
 * consumed `action`, as the reference appends them after env.step).
 * gg_export_successful: compacts the envs with ever[e] != 0, in env order, into out (n_selected, horizon, F) -- episode-major,
 * the order LeRobotDataset.add_frame / save_episode receive them -- and writes their env ids and the count (DEVICE int32).
 * out / out_env_ids must have room for n_envs episodes / ids. */
int gg_frame_dim(const GgSim* sim);
int gg_record(GgSim* sim, float* traj, int32_t* ever, int t, int horizon, const float* action, const float* obs_agent,
              const float* obs_env, const float* reward, gg_stream_t stream);
int gg_export_successful(GgSim* sim, const float* traj, const int32_t* ever, int horizon, float* out, int32_t* out_env_ids,
                         int32_t* out_count, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-object scenes: `FrankaCubeStackEmptyBatch` (gym_genesis/tasks/franka/cube_stack.py:27-240) and the SO-101
 * `CubeStackBatch` (gym_genesis/tasks/so101/cube_stack_batch.py:25-226, scene tasks/utils.py:593-794).  The robot, the box
 * geometry / inertia shared by all cubes and the solver constants come from GgModel; GgStackTask adds the object count,
 * the per-object spawn ranges and the stacking thresholds.  Object 0 = cube_1 (picked), object 1 = cube_2 (target), the
 * rest are distractors.  Cube-cube box-box contacts are generated (RigidOptions(box_box_detection=True), :46).
 *
 * State buffer (caller-owned, float32, gg_stack_state_rows() x gg_stack_state_stride(), env index fastest):
 *   rows [0, nu) q | [nu, 2nu) qd | 13 rows per object (pos3 quat4 linvel3 angvel3) | nu + 6 n_obj rows warm start.
 * One gg_stack_step = four launches on `stream`: stack_prep_kernel (thread per env: masked reset, PD, CRBA, Cholesky,
 * RNEA, joint-limit rows, body-level broad phase, robot-block solve when no pad can touch anything), stack_cubes_kernel
 * (thread per (cube, env): register solves of the cubes that can touch nothing but the plane), stack_solve_kernel (warp per
 * env with bodies in reach of each other: narrow phase incl. cube-cube, Newton on the coupled dofs), stack_finish_kernel
 * (thread per env: integration, observations, reward / flags, counters, statistics).
 * ------------------------------------------------------------------------------------------------------------------ */
#define GG_MAX_OBJ 5
#define GG_STACK_MAX_CONTACTS 96
#define GG_STACK_ENV_STATE_DIM 14 /* cube1 pos3 + quat4 + (eef - cube1)3 + |.|1 + cube2 pos3   cube_stack.py:209 */

typedef struct GgStackTask {
    int32_t n_obj, max_contacts;
    float spawn_lo[GG_MAX_OBJ][2], spawn_hi[GG_MAX_OBJ][2];   /* per-object uniform xy ranges   cube_stack.py:125-142 */
    float stack_xy, stack_z;                                  /* reward thresholds              cube_stack.py:195     */
    float min_dist;      /* > 0: objects 0 and 1 are re-drawn until at least this far apart (so101/cube_stack_batch.py:70-86) */
    int32_t agent_qpos;  /* 0: agent_pos = eef pos3 + quat4 + gripper dofs (franka/cube_stack.py:211);
                            1: agent_pos = joint positions, n_dof floats (so101/cube_stack_batch.py:176)                    */
} GgStackTask;

typedef struct GgStackSim GgStackSim;

int gg_stack_create(const GgModel* model, const GgStackTask* task, int n_envs, int device, GgStackSim** out);
int gg_stack_destroy(GgStackSim* sim);
int gg_stack_configure(GgStackSim* sim, const GgConfig* cfg);
int gg_stack_state_rows(const GgStackSim* sim);
int gg_stack_state_stride(const GgStackSim* sim);
int gg_stack_bind_state(GgStackSim* sim, float* state, int32_t* step_count, int32_t* episode);
/* action (n_envs, n_dof); obs_agent (n_envs, gg_stack_agent_dim()); obs_env (n_envs, 14); reward / flags (n_envs); reset_mask nullable. */
int gg_stack_agent_dim(const GgStackSim* sim);
int gg_stack_step(GgStackSim* sim, const float* action, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated,
                  uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream);
/* Host-resident I/O, same contract as gg_step_host: the prep kernel reads the pinned (n_envs, n_dof) action batch in place over
 * PCIe / C2C, the finish kernel writes reward / terminated / truncated into the nullable pinned mirrors; wait != 0 drains the stream. */
int gg_stack_step_host(GgStackSim* sim, const GgHostIO* io, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated,
                       uint8_t* truncated, const uint8_t* reset_mask, gg_stream_t stream);
/* obj_xy (nullable): (n_envs, n_obj, 2) explicit spawn positions (the reference's host RandomState stream, drawn per object:
 * all x then all y); NULL -> device Philox draw keyed by (cfg.seed, global env id, episode, object). */
int gg_stack_reset(GgStackSim* sim, const uint8_t* mask, const float* obj_xy, float* obs_agent, float* obs_env, gg_stream_t stream);
/* env-major copies: q (n_envs, n_dof + 7 n_obj), qd and qacc (n_envs, n_dof + 6 n_obj). */
int gg_stack_get_state(GgStackSim* sim, float* q, float* qd, float* qacc, gg_stream_t stream);
int gg_stack_set_state(GgStackSim* sim, const float* q, const float* qd, const float* qacc, gg_stream_t stream);
int gg_stack_observe(GgStackSim* sim, float* obs_agent, float* obs_env, float* reward, uint8_t* terminated, gg_stream_t stream);
int gg_stack_stats(GgStackSim* sim, double* out8, int clear, gg_stream_t stream);
/* trajectory store of the stacking collectors (examples/franka/stack_cube_state.py, examples/so_101/collect_task_stack_cube_batch.py):
 * same contract as gg_record / gg_export_successful with F = gg_stack_frame_dim() = agent_dim + 14 + n_dof + 1 */
int gg_stack_frame_dim(const GgStackSim* sim);
int gg_stack_record(GgStackSim* sim, float* traj, int32_t* ever, int t, int horizon, const float* action, const float* obs_agent,
                    const float* obs_env, const float* reward, gg_stream_t stream);
int gg_stack_export_successful(GgStackSim* sim, const float* traj, const int32_t* ever, int horizon, float* out, int32_t* out_env_ids,
                               int32_t* out_count, gg_stream_t stream);
/* per-env diagnostics of the last step (DEVICE int32[n_envs] each, nullable): contacts, Newton iterations */
int gg_stack_info(GgStackSim* sim, int32_t* ncon, int32_t* iters, gg_stream_t stream);

/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t gg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GG_B200_H */
