/*
 * bbx_cuda.h -- the thin C-ABI CUDA layer underneath the Ballbox host library.
 *
 * Plain pointers and sizes only: no C++ types, no torch types.  The host
 * library (ballbox_b200/csrc/host/ballbox_host.c) is the only intended caller;
 * a maintainer of the reference binds the public API in ballbox.h instead (see
 * INTEGRATION.md).  Every function returns 0 on success or a non-zero code
 * (a cudaError_t, or BBX_ERR_*); bbx_dev_error() gives the sticky message.
 *
 * What each entry point replaces in the reference (ballbox.h file:line):
 *   bbx_dev_step        PhysicsStep and its seven stages            :1787-1817
 *                        ApplyForces + IntegrateVelocities           :1820-1925
 *                        CollectCollisions (all-pairs loops)         :1979-2115
 *                        sphere/sphere, sphere/box, SAT + manifold   :451-471, :1565-1626, :1040-1132, :2550-2843
 *                        SDF probing                                 :2074-2114, :2318-2331, :2413-2477
 *                        SolveConstraints (generate, presolve, GS)   :1176-1226, :1295-1489
 *                        IntegratePositions, UpdateSleepStates       :1928-1977, :1696-1784
 *   bbx_dev_upload      the host SoA arrays of SphereSystem/BoxSystem :62-95 -> HBM
 *   bbx_dev_download    HBM -> the same host arrays
 *   bbx_dev_download_constraints   world->constraints as left by the step :1176-1211
 */
#ifndef BBX_CUDA_H
#define BBX_CUDA_H

#include "ballbox.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct BbxDev BbxDev;

#define BBX_ERR_NO_DEVICE     10001
#define BBX_ERR_BAD_ARGUMENT  10002
#define BBX_ERR_OVERFLOW      10003
#define BBX_ERR_HOST_SDF      10004
#define BBX_ERR_STICKY        10005

/* Host arrays of n_worlds worlds laid out back to back with a fixed stride:
 * world w owns records [w*cap_s, w*cap_s + s_count[w]) of every sphere array
 * and [w*cap_b, w*cap_b + b_count[w]) of every box array.  A single world is
 * n_worlds = 1.  Vec3 arrays are 3 floats per record, Quat arrays 4. */
typedef struct {
    int n_worlds, cap_s, cap_b;
    const int* s_count;
    const int* b_count;
    float *s_pos, *s_radius, *s_vel, *s_force, *s_mass, *s_angvel, *s_torque, *s_inertia, *s_timer;
    unsigned char *s_static, *s_sleep, *s_hascb;     /* s_hascb may be NULL */
    float *b_pos, *b_size, *b_rot, *b_vel, *b_force, *b_mass, *b_angvel, *b_torque, *b_inertia, *b_timer;
    unsigned char *b_static, *b_sleep, *b_hascb;     /* b_hascb may be NULL */
} BbxHostBodies;

typedef struct {
    float gravity[3];
    float dt;
    PhysicsSettings settings;      /* by value, re-read at every step like the reference */
    int   sdf_id;                  /* BBX_SDF_* of ballbox_ext.h; 0 = no SDF              */
    float sdf_params[8];
    int   solver_schedule;         /* BBX_SOLVER_EXACT / _COLOURED / _EXACT_FLOW           */
    int   has_forces;              /* 0: force/torque arrays are known to be all zero      */
    int   want_callbacks;          /* collect callback events                              */
    float warm_start;              /* 0: impulses start at zero like the reference (:1203-1206); f > 0: a contact that
                                      existed in the previous step starts from f x its final impulses (ballbox_ext.h) */
} BbxStepParams;

/* baked SDF: nx*ny*nz samples (x fastest) of a host function at origin + (i,j,k)*cell; replaces any earlier bake.
 * channels = 1: the distance; channels = 4: distance followed by the three central differences of
 * SDF_EstimateNormal (ballbox.h:2325-2327) at the node, which the kernels interpolate instead of differencing the
 * interpolated distance */
int  bbx_dev_set_baked_sdf(BbxDev* dev, const float* values, int channels, int nx, int ny, int nz, const float origin[3], float cell);

int  bbx_device_count(void);
int  bbx_dev_create(BbxDev** out, int device, int n_worlds, int cap_spheres, int cap_boxes,
                    int cap_contacts_per_world);
void bbx_dev_destroy(BbxDev* dev);
void bbx_dev_set_stream(BbxDev* dev, void* cuda_stream);
int  bbx_dev_upload(BbxDev* dev, const BbxHostBodies* hb);
/* forces and torques only (AddSphereForce & co. between resident steps) */
int  bbx_dev_upload_forces(BbxDev* dev, const BbxHostBodies* hb);
/* body slots (sphere (w,i) = w*cap_s + i, box (w,i) = n_worlds*cap_s + w*cap_b + i) that Add*Force/Torque touched since the
 * last upload: a sleeping one wakes up, timer 0 (ballbox.h:420-423, :434-437, :793-796, :807-810), even for a zero force */
int  bbx_dev_wake_bodies(BbxDev* dev, const int* slots, int n);
int  bbx_dev_step(BbxDev* dev, const BbxStepParams* params, int steps);
int  bbx_dev_download(BbxDev* dev, const BbxHostBodies* hb);
/* Overlap helpers of the drop-in PhysicsStep (host arrays in, host arrays out):
 *   bbx_dev_wait_upload     returns when the last bbx_dev_upload has finished READING the host arrays (the step
 *                           itself may still be running): the caller may then overwrite them (force clearing :1836)
 *   bbx_dev_download_begin  like bbx_dev_download, but the copies are only queued (on a second stream)
 *   bbx_dev_download_end    waits for them and reports a contact overflow like bbx_dev_download
 * bbx_dev_download_constraints between begin and end shares the copy stream: body arrays travel while the
 * constraint records are still being exported and sorted. */
int  bbx_dev_wait_upload(BbxDev* dev);
int  bbx_dev_download_begin(BbxDev* dev, const BbxHostBodies* hb);
int  bbx_dev_download_end(BbxDev* dev);

/* ---- the reference's public stage functions (ballbox.h:217-229), single world only -----------
 * bbx_dev_run_stage runs ONE stage on the state uploaded before it:
 *   APPLY_FORCES   ApplyForces :1820           INTEGRATE_V  IntegrateVelocities :1841
 *   COLLECT        CollectCollisions :1979     PRESOLVE     PreSolveConstraints :1295 (imported list)
 *   SOLVE_SWEEP    SolveVelocityConstraints :1410, one sweep (imported list with solver fields)
 *   SOLVE_ALL      SolveConstraints :1213 = generate + presolve + solver_iterations sweeps (imported contacts)
 *   INTEGRATE_P    IntegratePositions :1928    SLEEP        UpdateSleepStates :1696
 * Host contact lists are imported in ARRAY order, whatever the user did to them between stages. */
#define BBX_STAGE_APPLY_FORCES 1
#define BBX_STAGE_INTEGRATE_V  2
#define BBX_STAGE_COLLECT      3
#define BBX_STAGE_PRESOLVE     4
#define BBX_STAGE_SOLVE_SWEEP  5
#define BBX_STAGE_SOLVE_ALL    6
#define BBX_STAGE_INTEGRATE_P  7
#define BBX_STAGE_SLEEP        8
int  bbx_dev_run_stage(BbxDev* dev, const BbxStepParams* params, int stage);
int  bbx_dev_download_forces(BbxDev* dev, const BbxHostBodies* hb);
/* after COLLECT: contacts in emission order (44-byte records) */
int  bbx_dev_download_contacts(BbxDev* dev, CollisionContact* out, int cap, int* count);
/* host list -> device (array order).  Exactly one of contacts / constraints is non-NULL. */
int  bbx_dev_import_list(BbxDev* dev, const CollisionContact* contacts, const ContactConstraint* constraints, int n);
/* device list -> host records in array order (after PRESOLVE / SOLVE_SWEEP / SOLVE_ALL) */
int  bbx_dev_export_list(BbxDev* dev, ContactConstraint* out, int n, float restitution, float friction, int solved);
/* per world: out + w*cap_per_world receives that world's constraints sorted
 * into the reference's emission order; counts[w] their number. */
int  bbx_dev_download_constraints(BbxDev* dev, ContactConstraint* out, int cap_per_world, int* counts);
/* Collision-callback events (reference ballbox.h:1999-2005, :2034-2044, :2063-2069, :2100-2103):
 * while params->want_callbacks is set, every step appends one record per contact (per PAIR for
 * box-box) that involves a body with a registered callback to a device buffer.
 * bbx_dev_fetch_events copies them out sorted by (step, emission order) and empties the buffer:
 * contacts[e] is the contact, steps[e] the step index since the last fetch, worlds[e] the world.
 * Returns BBX_ERR_OVERFLOW if more than the buffer capacity were produced (the first `capacity`
 * in arrival order are kept). */
int  bbx_dev_event_capacity(BbxDev* dev);
int  bbx_dev_fetch_events(BbxDev* dev, CollisionContact* contacts, int* steps, int* worlds, int max_events, int* count);
/* refresh the has-callback bits after Set*CollisionCallback on a resident world */
int  bbx_dev_upload_callback_flags(BbxDev* dev, const BbxHostBodies* hb);
int  bbx_dev_synchronize(BbxDev* dev);
int  bbx_dev_stats(BbxDev* dev, BallboxStats* out);
/* per-stage device times: when on, every step records CUDA events on the step stream.
 * bbx_dev_get_timing sums them (ms) into out6 = {integrate-v, broadphase, narrowphase,
 * presolve+graph, solve kernel, integrate-p+sleep}, reports the number of steps covered
 * and resets the accumulation. */
void bbx_dev_set_timing(BbxDev* dev, int on);
int  bbx_dev_get_timing(BbxDev* dev, float* out6, int* steps);
/* the staged replay kernel alone (part of out6[4]), for the steps covered by the last bbx_dev_get_timing call */
int  bbx_dev_get_replay_timing(BbxDev* dev, float* ms, int* steps);
/* debug trace of the last step: node counts per level and class (8 ints per level), and
 * ns[l+1] = globaltimer at the end of level l of solver iteration 1 (ns[l+1]-ns[l] = level time, l >= 1) */
int  bbx_dev_debug_levels(BbxDev* dev, int* counts, unsigned long long* ns, int max_levels);
int  bbx_dev_debug_warp_trace(BbxDev* dev, unsigned long long* out, int n_warps);
int  bbx_dev_debug_schedule(BbxDev* dev, int* out4, int cap_nodes);   /* nodes of the last level-path solve in visiting order */
int  bbx_dev_debug_deferred(BbxDev* dev);        /* call after bbx_dev_stats */
const char* bbx_dev_error(BbxDev* dev);          /* NULL when healthy */
void bbx_dev_clear_error(BbxDev* dev);

/* ---- zero-copy access for batched / RL use ------------------
 * Raw device pointers of the working state.  N = n_worlds*(cap_s+cap_b) body slots, spheres
 * first: sphere (w,i) = w*cap_s + i, box (w,i) = n_worlds*cap_s + w*cap_b + i. */
typedef struct {
    int   n_worlds, cap_spheres, cap_boxes;
    void* pos;          /* float4[N]   xyz, bounding radius                                  */
    void* vel;          /* float4[2N]  [2g] = (v, 1/m), [2g+1] = (w, sphere 1/I)             */
    void* box_rot;      /* float4[2NB] [2b] = rotation, [2b+1] = 1/diag(I)                   */
    void* flags;        /* uint8[N]    1 static, 2 sleeping, 4 valid, 8 callback, 16 big     */
    void* sphere_force; /* float[3NS]  force accumulators, consumed and cleared by a step    */
    void* sphere_torque;
    void* box_force;    /* float[3NB] */
    void* box_torque;
} BbxDeviceViews;
int  bbx_dev_views(BbxDev* dev, BbxDeviceViews* out);
/* accumulators += the given DEVICE arrays (any may be NULL); bodies that receive a non-zero
 * force or torque are woken like AddSphereForce does (ballbox.h:420-423) */
int  bbx_dev_add_wrench(BbxDev* dev, const float* d_sphere_force, const float* d_sphere_torque,
                        const float* d_box_force, const float* d_box_torque);
/* keep a device copy of the current state / restore the listed worlds from it */
int  bbx_dev_snapshot(BbxDev* dev);
int  bbx_dev_restore_worlds(BbxDev* dev, const int* world_ids, int n);

/* ---- joints and springs (extension, ballbox_ext.h BallboxAddSpring / BallboxAdd*Joint; sequential definition:
 * oracle/bbx_oracle.c).  Bodies are global slots: sphere (w,i) = w*cap_s + i, box (w,i) = n_worlds*cap_s + w*cap_b + i,
 * -1 = the world (end B only).  Anchors are in the body frame (spheres: the centre, 0; the world: the point itself).
 * A call replaces the whole list; lists are in creation order, which is the order the sequential definition sweeps
 * joints and accumulates spring forces in. */
#define BBX_JOINT_BALL 0   /* three rows (world x, y, z) holding the two anchor points together */
#define BBX_JOINT_ROD  1   /* one row along the line between the anchor points keeping their distance at `rest` */
typedef struct { int kind, ga, gb; float la[3], lb[3]; float rest; } BbxJointDef;
typedef struct { int ga, gb; float la[3], lb[3]; float rest, k, c; } BbxSpringDef;
int  bbx_dev_set_joints(BbxDev* dev, const BbxJointDef* joints, int n);
int  bbx_dev_set_springs(BbxDev* dev, const BbxSpringDef* springs, int n);

/* pinned host memory for the public arrays (falls back to calloc without a device) */
void* bbx_host_alloc(size_t bytes, int* pinned);
void  bbx_host_free(void* p, int pinned);

#ifdef __cplusplus
}
#endif
#endif /* BBX_CUDA_H */
