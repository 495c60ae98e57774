/*
 * ballbox_ext.h -- extensions of the B200-native library that the reference
 * does not have.  Nothing here changes the meaning of the reference API in
 * ballbox.h; a program that never calls these gets the drop-in behaviour
 * (BBX_SYNC_COMPAT: host arrays are uploaded before and downloaded after every
 * PhysicsStep, exactly as if the step had run on the host).
 */
#ifndef BALLBOX_EXT_H
#define BALLBOX_EXT_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- residency ------------------------------------------------------------ */
#define BBX_SYNC_COMPAT    0   /* default: upload / step / download in every PhysicsStep   */
#define BBX_SYNC_RESIDENT  1   /* state stays in HBM; host arrays move only on request     */
#define BBX_SYNC_COMPAT_BODIES 2 /* like COMPAT, but world->constraints stays on the device
                                  (count 0) until BallboxDownload(world, BBX_DL_CONSTRAINTS):
                                  the list is ~6x the body state; callbacks still fire        */

#define BBX_DL_BODIES       1  /* positions, rotations, velocities, sleep state            */
#define BBX_DL_CONSTRAINTS  2  /* world->constraints in the reference's emission order     */

void BallboxSetSyncMode(PhysicsWorld* world, int mode);
int  BallboxUpload(PhysicsWorld* world);               /* host arrays -> device; 0 = ok       */
int  BallboxDownload(PhysicsWorld* world, int what);   /* device -> host arrays; 0 = ok       */
int  PhysicsStepN(PhysicsWorld* world, float deltaTime, int steps);  /* resident stepping     */
int  BallboxSynchronize(PhysicsWorld* world);          /* wait for queued device work         */

/* ---- device selection; `stream` is a cudaStream_t passed as void* ---------- */
int  BallboxSetDevice(PhysicsWorld* world, int device);
void BallboxSetStream(PhysicsWorld* world, void* stream);

/* ---- error channel (the reference has none: a failed CUDA call or a buffer
 * overflow makes PhysicsStep a no-op and leaves a sticky message here) ------- */
const char* BallboxGetLastError(PhysicsWorld* world);  /* NULL when healthy */
void        BallboxClearError(PhysicsWorld* world);

/* ---- SDF colliders on the device.  A host function pointer cannot run on the GPU.
 * UseSDF(world, fn) (reference ballbox.h:1688) keeps the pointer; at the next PhysicsStep the
 * function is BAKED ONCE: sampled on a regular grid over the bake region and uploaded, and
 * the kernels read it back with trilinear interpolation (f32, fixed operation order).  A
 * baked SDF is an approximation: the error is the interpolation error of fn at the grid
 * spacing (O(h^2 |f''|); exact for functions that are linear in x, y and z); normals interpolate the
 * reference's central differences sampled at the grid nodes (BallboxSetSDFBakeGradient; accuracy table:
 * profiles/r01_sdf_bake.md).  Outside the region the nearest grid
 * value is used.  For bit-exact results select an SDF that exists as device code with
 * UseSDFDevice instead (it takes precedence over a host pointer). ---------------------- */
#define BBX_SDF_NONE        0
#define BBX_SDF_PLANE_Y     1  /* f(p) = p.y - params[0]                                    */
#define BBX_SDF_WAVY_DET    2  /* f(p) = p.y - S(p.x)*S(p.z), S = bbx_det_sin (bit-exact on
                                  CPU and GPU; see bbx_sdf_library.h)                       */
#define BBX_SDF_WAVY_LIBM   3  /* f(p) = p.y - sinf(p.x)*sinf(p.z) with CUDA sinf (README)  */
#define BBX_SDF_SPHERE_IN   4  /* inside of a sphere: f(p) = params[0] - |p|                */
#define BBX_SDF_BAKED       5  /* reported by BallboxSdfInUse when a host function was baked */
void UseSDFDevice(PhysicsWorld* world, int sdf_id, const float* params8);
/* Region and resolution of the bake (cells along the longest axis, 2..1024; the other axes get the
 * same spacing).  Default: the bounding box of all bodies at bake time, grown by 25 % + 2 units on
 * every side, 192 cells along the longest axis.  Calling it (or UseSDF with another function)
 * schedules a new bake. */
void BallboxSetSDFBakeRegion(PhysicsWorld* world, Vec3 lo, Vec3 hi, int resolution);
/* Gradient channels (default on): the bake also samples the three central differences of SDF_EstimateNormal
 * (ballbox.h:2318-2331, epsilon = settings.sdf_normal_epsilon at bake time; a later change of that setting re-bakes)
 * and contact normals interpolate THEM, error O(h^2), instead of differencing the interpolated distance, error O(h).
 * Costs 4x the grid memory and 7 calls of fn per node.  0 turns them off. */
void BallboxSetSDFBakeGradient(PhysicsWorld* world, int on);
/* Bake now instead of at the next step; returns 0 or an error code (message: BallboxGetLastError). */
int  BallboxBakeSDF(PhysicsWorld* world);
int  BallboxSdfInUse(PhysicsWorld* world);      /* BBX_SDF_* the next step will use */

/* ---- solver schedule ------------------------------------------------------- */
#define BBX_SOLVER_EXACT     0 /* level-scheduled Gauss-Seidel: bit-identical to the
                                  reference's sequential sweep (default)                    */
#define BBX_SOLVER_COLOURED  1 /* graph-coloured Gauss-Seidel (greedy colours, ~max degree + 1 classes): reorders the sweep */
#define BBX_SOLVER_EXACT_FLOW 2 /* the same exact schedule, barrier-free: every constraint waits
                                  only for its own two predecessors (versioned body records,
                                  iterations overlap); bit-identical results, experimental:
                                  measured slower than the level kernel at 1M bodies
                                  (profiles/r01_notes.md)                                   */
void BallboxSetSolverSchedule(PhysicsWorld* world, int schedule);

/* ---- warm starting (NOT in the reference: its ContactConstraint reserves the accumulated impulses "for warm
 * starting" but GenerateContactConstraints zeroes them, ballbox.h:134-136, :1203-1206).  factor = 0 (default) keeps
 * the reference's behaviour.  With factor > 0 a contact that also existed in the previous step -- same body A, same
 * body B, same index k among the contacts of that pair in emission order -- starts from factor x the impulses it
 * ended the previous step with, and, when solver_iterations > 0, a pass before the first sweep applies them:
 * for every matched constraint in array order, A receives -n*jn and B +n*jn, then (friction > 0) the same for
 * t1*jt0 and t2*jt1, through ApplyConstraintImpulse.  Unmatched contacts start from zero and apply nothing.
 * The sequential definition is restated in oracle/bbx_oracle.c (OracleSetWarmStarting) and the GPU path is
 * bit-identical to it with the exact schedules.  PhysicsStep / PhysicsStepN / PhysicsStepBatch only; the public
 * stage functions ignore it.  A world reset (BatchResetWorlds) forgets that world's impulses. */
void BallboxSetWarmStarting(PhysicsWorld* world, float factor);

/* ---- joints and springs (NOT in the reference: its README lists "constraints, like hinges, springs" as planned,
 * README.md:25).  A body is (type, index): type 0 = sphere, 1 = box, -1 = the world (a fixed point; end B only).
 * Anchors are given in world coordinates at the time of the call and stored in the body's frame.  A sphere has no
 * orientation in this engine (SphereSystem has no rotations), so a sphere is always attached at its CENTRE: the
 * anchor given for a sphere end is replaced by the centre.  All calls return the index of the new element among the
 * world's joints / springs, or -1.
 *   spring          force k (len - rest) + c (relative anchor velocity along the line), added to the force / torque
 *                   accumulators of both bodies in creation order BEFORE gravity (ApplyForces); wakes dynamic ends
 *   ball joint      three bilateral velocity rows (world x, y, z) that hold the two anchor points together
 *   distance joint  one bilateral row along the line between the anchors that keeps the distance they had when the
 *                   joint was created (a rigid rod; with a sphere end this is the pendulum)
 *   hinge           two ball joints at anchor -+ half_span * axis (returns the index of the first)
 * Joint rows are presolved after the contacts (bias = -baumgarte_bias / dt * position error) and swept after the contacts
 * in every solver iteration, joints in creation order; they use ApplyConstraintImpulse, so jointed dynamic bodies stay
 * awake.  The sequential definition is restated in oracle/bbx_oracle.c and the GPU path is bit-identical to it.  A
 * world with joints always runs the exact level schedule (no coloured / flow / group solver), and its contact graph may
 * be at most 256 levels deep (a deeper one is reported as an overflow error).  PhysicsStep / PhysicsStepN /
 * PhysicsStepBatch only; the public stage functions ignore joints and springs.  In a RESIDENT world the call first
 * downloads the bodies so that anchors refer to the current positions. */
int  BallboxAddSpring(PhysicsWorld* world, int typeA, int indexA, int typeB, int indexB, Vec3 anchorA, Vec3 anchorB,
                      float rest_length, float stiffness, float damping);
int  BallboxAddBallJoint(PhysicsWorld* world, int typeA, int indexA, int typeB, int indexB, Vec3 anchor);
int  BallboxAddDistanceJoint(PhysicsWorld* world, int typeA, int indexA, int typeB, int indexB, Vec3 anchorA, Vec3 anchorB);
int  BallboxAddHingeJoint(PhysicsWorld* world, int typeA, int indexA, int typeB, int indexB, Vec3 anchor, Vec3 axis, float half_span);
int  BallboxJointCount(PhysicsWorld* world);
int  BallboxSpringCount(PhysicsWorld* world);
void BallboxClearJoints(PhysicsWorld* world);          /* removes the world's joints AND springs */

/* ---- statistics of the most recent step (and totals since creation) -------- */
typedef struct {
    long long steps;               /* device steps executed                                 */
    long long body_steps;          /* sum over steps of (spheres + boxes)                   */
    long long kernel_launches;     /* CUDA kernels launched by this library                 */
    int  spheres, boxes;           /* bodies on the device                                  */
    int  candidate_pairs;          /* broadphase output of the last step (SS+SB+BB)         */
    int  contacts;                 /* contacts of the last step                             */
    int  solver_contacts;          /* contacts touching >= 1 dynamic body                   */
    int  touched_bodies;           /* dynamic bodies with >= 1 contact                      */
    int  levels;                   /* dependency levels (exact) or colours (coloured)       */
    int  overflow;                 /* bit0 pairs, bit1 contacts, bit2 grid, bit3 callbacks  */
    int  grid_cells;
    int  reserved[5];
} BallboxStats;
int BallboxGetStats(PhysicsWorld* world, BallboxStats* out);

/* ---- per-stage device timing (CUDA events on the step stream).  out6 (ms, summed
 * over the steps since the last query) = integrate-v, broadphase, narrowphase,
 * presolve+graph, solve kernel, integrate-p+sleep. ------------------------------ */
void BallboxSetStageTiming(PhysicsWorld* world, int on);
int  BallboxGetStageTiming(PhysicsWorld* world, float* out6, int* steps);
int  BallboxGetReplayTiming(PhysicsWorld* world, float* ms, int* steps);      /* k_replay_staged alone (part of out6[4]); call after BallboxGetStageTiming */

int  BallboxDebugLevels(PhysicsWorld* world, int* counts8, unsigned long long* ns, int max_levels);
/* visiting order of the last solve (single world, level path): out4[4 i ..] = (level or colour, constraints of the node,
   dynamic body A or -1, dynamic body B or -1); returns the number of nodes.  Tests use it to check that a level is an
   independent set. */
int  BallboxDebugSchedule(PhysicsWorld* world, int* out4, int cap_nodes);
int  BallboxDebugDeferred(PhysicsWorld* world);   /* after BallboxGetStats: heavy terminal nodes the last step swept in its extra final level */
int  BallboxDebugWarpTrace(PhysicsWorld* world, unsigned long long* out4, int n_warps);   /* BBX_TRACE_LEVEL=l+1: per warp of the staged replay in level l, iteration 1: start, data ready, end (ns), constraints<<16|chunks */

/* ---- batches of independent worlds (RL style).  All worlds of a batch share
 * one device context and are stepped by the same kernels; worlds never
 * interact.  Settings, gravity and SDF are taken from world 0. --------------- */
typedef struct PhysicsWorldBatch PhysicsWorldBatch;
PhysicsWorldBatch* CreatePhysicsWorldBatch(int n_worlds, int max_spheres, int max_boxes,
                                           int max_collisions);
void          DestroyPhysicsWorldBatch(PhysicsWorldBatch* batch);
int           BatchWorldCount(PhysicsWorldBatch* batch);
PhysicsWorld* BatchWorld(PhysicsWorldBatch* batch, int index);
int  BatchSetDevice(PhysicsWorldBatch* batch, int device);
void BatchSetStream(PhysicsWorldBatch* batch, void* stream);
int  BatchUpload(PhysicsWorldBatch* batch);
int  BatchDownload(PhysicsWorldBatch* batch, int what);
int  PhysicsStepBatch(PhysicsWorldBatch* batch, float deltaTime, int steps);
int  BatchSynchronize(PhysicsWorldBatch* batch);
int  BatchGetStats(PhysicsWorldBatch* batch, BallboxStats* out);
void BatchSetStageTiming(PhysicsWorldBatch* batch, int on);
int  BatchGetStageTiming(PhysicsWorldBatch* batch, float* out6, int* steps);
const char* BatchGetLastError(PhysicsWorldBatch* batch);

/* ---- zero-copy batch interface (RL style).  BatchGetDeviceViews exposes the device arrays
 * (layout documented at BbxDeviceViews in bbx_cuda.h) so that observations can be read and
 * forces written without touching the host; BatchAddWrenchDevice adds per-body force/torque
 * arrays that already live on the device; BatchSnapshot / BatchResetWorlds implement episode
 * resets on the device. ------------------------------------------------------------------- */
typedef struct {
    int   n_worlds, cap_spheres, cap_boxes;
    void *pos, *vel, *box_rot, *flags, *sphere_force, *sphere_torque, *box_force, *box_torque;
} BallboxDeviceViews;
int BatchGetDeviceViews(PhysicsWorldBatch* batch, BallboxDeviceViews* out);
int BatchAddWrenchDevice(PhysicsWorldBatch* batch, const float* d_sphere_force, const float* d_sphere_torque,
                         const float* d_box_force, const float* d_box_torque);
int BatchSnapshot(PhysicsWorldBatch* batch);
int BatchResetWorlds(PhysicsWorldBatch* batch, const int* world_ids, int n);

/* ---- library identity ------------------------------------------------------ */
const char* BallboxBackendName(void);   /* "ballbox-b200/cuda-sm_100a" */
int         BallboxDeviceCount(void);   /* 0 when no CUDA device is visible */

#ifdef __cplusplus
}
#endif
#endif /* BALLBOX_EXT_H */
