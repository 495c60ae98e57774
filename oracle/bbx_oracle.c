/*
 * bbx_oracle.c -- TEST INFRASTRUCTURE: a plain-C, single-threaded restatement of
 * the reference's PhysicsStep (reference ballbox.h:1787-1817 and everything it
 * calls).  It is the checker the CUDA path is compared with, bit for bit.  The
 * product never includes, links or calls this file; only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs load the resulting
 * oracle/libbbx_oracle.so.
 *
 * PINNING.  The reference ships no tests or golden vectors,
 * so this restatement is pinned against the reference itself: oracle/_ref is the
 * unmodified header compiled here, tests/test_oracle.py checks that this file
 * reproduces it bit-for-bit (state and constraint arrays, every step) on all five
 * scene shapes, and tests/golden/ holds vectors generated from oracle/_ref.
 *
 * Build: gcc -O2 -std=gnu11 -ffp-contract=off (no FMA contraction).
 *
 * Structure differs from the reference on purpose (it is a restatement, not a
 * copy): contacts are emitted through one `emit` helper, detection can run behind
 * a pruning grid (OraclePhysicsStepPruned), and the box-box path works on an Obb
 * value type.  The ARITHMETIC (operation order, comparisons, constants) follows
 * the cited reference lines exactly.
 */
#include "ballbox.h"
#include "../ballbox_b200/csrc/host/bbx_scenes.c"
#include "prune_grid.h"

/* ===== math: ballbox.h:501-639 ========================================== */
Vec3 Vec3Add(Vec3 a, Vec3 b) { return (Vec3){ a.x + b.x, a.y + b.y, a.z + b.z }; }
Vec3 Vec3Sub(Vec3 a, Vec3 b) { return (Vec3){ a.x - b.x, a.y - b.y, a.z - b.z }; }
Vec3 Vec3Scale(Vec3 v, float s) { return (Vec3){ v.x * s, v.y * s, v.z * s }; }
float Vec3Dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
Vec3 Vec3Cross(Vec3 a, Vec3 b) { return (Vec3){ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
float Vec3Length(Vec3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
Vec3 Vec3Normalize(Vec3 v) { float l = Vec3Length(v); if (l < 1e-6f) return (Vec3){0, 0, 0}; return Vec3Scale(v, 1.0f / l); }
Quat QuatIdentity(void) { return (Quat){0.0f, 0.0f, 0.0f, 1.0f}; }
Quat QuatMultiply(Quat p, Quat q)
{
    return (Quat){ p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
                   p.w * q.y - p.x * q.z + p.y * q.w + p.z * q.x,
                   p.w * q.z + p.x * q.y - p.y * q.x + p.z * q.w,
                   p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z };
}
Quat QuatNormalize(Quat q)
{
    float l = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    if (l == 0.0f) return QuatIdentity();
    return (Quat){ q.x / l, q.y / l, q.z / l, q.w / l };
}
Quat QuatConjugate(Quat q) { return (Quat){ -q.x, -q.y, -q.z, q.w }; }
Quat QuatFromAxisAngle(Vec3 axis, float angle)          /* :564-575 */
{
    Vec3 n = Vec3Scale(axis, 1.0f / sqrtf(Vec3Dot(axis, axis)));
    float h = angle * 0.5f, s = sinf(h);
    return (Quat){ n.x * s, n.y * s, n.z * s, cosf(h) };
}
Vec3 QuatRotateVec3(Quat q, Vec3 v)                     /* :630-639 */
{
    Vec3 u = { q.x, q.y, q.z }; float s = q.w;
    return Vec3Add(Vec3Add(Vec3Scale(v, s * s - Vec3Dot(u, u)), Vec3Scale(u, 2.0f * Vec3Dot(u, v))),
                   Vec3Scale(Vec3Cross(u, v), 2.0f * s));
}

/* ===== world management: ballbox.h:301-445, :642-818, :1629-1693, :2483-2544 */
void ResetPhysicsSettingsToDefaults(PhysicsWorld* w)
{
    if (!w) return;
    PhysicsSettings* s = &w->settings;
    s->linear_damping = 0.98f; s->angular_damping = 0.98f; s->restitution = 0.0f; s->friction = 0.1f;
    s->solver_iterations = 25; s->baumgarte_bias = 0.3f; s->allowed_penetration = 0.05f; s->velocity_threshold = 0.02f;
    s->sleep_linear_threshold = 0.1f; s->sleep_angular_threshold = 0.1f; s->sleep_time_required = 0.05f;
    s->manifold_contact_tolerance = 0.01f; s->manifold_penetration_tolerance = -0.5f; s->manifold_max_contacts = 8;
    s->collision_epsilon = 1e-6f; s->sdf_normal_epsilon = 0.001f; s->vector_normalize_epsilon = 1e-6f; s->quaternion_epsilon = 1e-6f;
}
void SetDamping(PhysicsWorld* w, float l, float a) { if (w) { w->settings.linear_damping = l; w->settings.angular_damping = a; } }
void SetRestitution(PhysicsWorld* w, float r) { if (w) w->settings.restitution = r; }
void SetCorrectionParameters(PhysicsWorld* w, float b, float a) { if (w) { w->settings.baumgarte_bias = b; w->settings.allowed_penetration = a; } }
void SetSolverIterations(PhysicsWorld* w, int i) { if (w) w->settings.solver_iterations = i; }
void SetManifoldSettings(PhysicsWorld* w, float t, float p, int m)
{
    if (!w) return;
    w->settings.manifold_contact_tolerance = t; w->settings.manifold_penetration_tolerance = p;
    w->settings.manifold_max_contacts = m > MAX_MANIFOLD_CONTACTS ? MAX_MANIFOLD_CONTACTS : m;
}

#define NEW(T, n) ((T*)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(T)))
PhysicsWorld* CreatePhysicsWorld(int ms, int mb, int mc)
{
    PhysicsWorld* w = NEW(PhysicsWorld, 1);
    SphereSystem* s = w->spheres = NEW(SphereSystem, 1);
    BoxSystem* b = w->boxes = NEW(BoxSystem, 1);
    s->positions = NEW(Vec3, ms); s->radii = NEW(float, ms); s->velocities = NEW(Vec3, ms); s->forces = NEW(Vec3, ms);
    s->masses = NEW(float, ms); s->angular_velocities = NEW(Vec3, ms); s->torques = NEW(Vec3, ms); s->inertias = NEW(float, ms);
    s->is_static = NEW(bool, ms); s->is_sleeping = NEW(bool, ms); s->sleep_timers = NEW(float, ms);
    s->collision_callbacks = NEW(OnCollision, ms); s->capacity = ms;
    for (int i = 0; i < ms; i++) { s->radii[i] = 1.0f; s->masses[i] = 1.0f; s->inertias[i] = 1.0f; }
    b->positions = NEW(Vec3, mb); b->sizes = NEW(Vec3, mb); b->rotations = NEW(Quat, mb); b->velocities = NEW(Vec3, mb);
    b->forces = NEW(Vec3, mb); b->masses = NEW(float, mb); b->angular_velocities = NEW(Vec3, mb); b->torques = NEW(Vec3, mb);
    b->inertias = NEW(Vec3, mb); b->is_static = NEW(bool, mb); b->is_sleeping = NEW(bool, mb); b->sleep_timers = NEW(float, mb);
    b->collision_callbacks = NEW(OnCollision, mb); b->capacity = mb;
    for (int i = 0; i < mb; i++) {
        b->sizes[i] = (Vec3){1, 1, 1}; b->rotations[i] = QuatIdentity(); b->masses[i] = 1.0f;
        b->inertias[i] = (Vec3){1.0f / 6.0f, 1.0f / 6.0f, 1.0f / 6.0f};
    }
    w->collisions = NEW(CollisionCollection, 1);
    w->collisions->contacts = NEW(CollisionContact, mc); w->collisions->capacity = mc;
    w->constraints = NEW(ConstraintCollection, 1);
    w->constraints->constraints = NEW(ContactConstraint, mc); w->constraints->capacity = mc;
    w->dt = 1.0f / 60.0f;
    ResetPhysicsSettingsToDefaults(w);
    return w;
}
static void warm_forget(PhysicsWorld* w);
static void joints_forget(PhysicsWorld* w);
void DestroyPhysicsWorld(PhysicsWorld* w)
{
    if (!w) return;
    warm_forget(w);
    joints_forget(w);
    SphereSystem* s = w->spheres; BoxSystem* b = w->boxes;
    free(s->positions); free(s->radii); free(s->velocities); free(s->forces); free(s->masses); free(s->angular_velocities);
    free(s->torques); free(s->inertias); free(s->is_static); free(s->is_sleeping); free(s->sleep_timers); free(s->collision_callbacks); free(s);
    free(b->positions); free(b->sizes); free(b->rotations); free(b->velocities); free(b->forces); free(b->masses);
    free(b->angular_velocities); free(b->torques); free(b->inertias); free(b->is_static); free(b->is_sleeping);
    free(b->sleep_timers); free(b->collision_callbacks); free(b);
    free(w->collisions->contacts); free(w->collisions); free(w->constraints->constraints); free(w->constraints); free(w);
}
void UseSDF(PhysicsWorld* w, SDFFunction f) { if (w) { w->SDFCollider = f; w->usesSDF = (f != NULL); } }

int AddSphere(PhysicsWorld* w, Vec3 p, float r)
{
    if (!w || !w->spheres || w->spheres->count >= w->spheres->capacity) return -1;
    SphereSystem* s = w->spheres; int i = s->count++;
    s->positions[i] = p; s->radii[i] = r; s->masses[i] = 1.0f;
    s->velocities[i] = s->forces[i] = s->angular_velocities[i] = s->torques[i] = (Vec3){0, 0, 0};
    s->inertias[i] = (2.0f / 5.0f) * s->masses[i] * r * r;                                   /* :376 */
    return i;
}
static Vec3 oracle_box_inertia(float m, Vec3 h)                                               /* :723-731 */
{
    float w = h.x * 2.0f, hh = h.y * 2.0f, d = h.z * 2.0f;
    return (Vec3){ (m / 12.0f) * (hh * hh + d * d), (m / 12.0f) * (w * w + d * d), (m / 12.0f) * (w * w + hh * hh) };
}
int AddBox(PhysicsWorld* w, Vec3 p, Vec3 half, Quat q)
{
    if (!w || !w->boxes || w->boxes->count >= w->boxes->capacity) return -1;
    BoxSystem* b = w->boxes; int i = b->count++;
    b->positions[i] = p; b->sizes[i] = half; b->rotations[i] = q; b->masses[i] = 1.0f;
    b->velocities[i] = b->forces[i] = b->angular_velocities[i] = b->torques[i] = (Vec3){0, 0, 0};
    b->inertias[i] = oracle_box_inertia(1.0f, half);
    b->is_static[i] = false; b->is_sleeping[i] = false; b->sleep_timers[i] = 0.0f;
    return i;
}
void SetSphereMass(PhysicsWorld* w, int i, float m)
{
    if (!w || i < 0 || i >= w->spheres->count || m <= 0.0f) return;
    w->spheres->masses[i] = m; float r = w->spheres->radii[i]; w->spheres->inertias[i] = (2.0f / 5.0f) * m * r * r;
}
void SetBoxMass(PhysicsWorld* w, int i, float m)
{
    if (!w || i < 0 || i >= w->boxes->count || m <= 0.0f) return;
    w->boxes->masses[i] = m; w->boxes->inertias[i] = oracle_box_inertia(m, w->boxes->sizes[i]);
}
void SetSphereStatic(PhysicsWorld* w, int i, bool st)
{
    if (!w || i < 0 || i >= w->spheres->count) return;
    SphereSystem* s = w->spheres; s->is_static[i] = st;
    if (st) s->velocities[i] = s->angular_velocities[i] = s->forces[i] = s->torques[i] = (Vec3){0, 0, 0};
}
void SetBoxStatic(PhysicsWorld* w, int i, bool st)
{
    if (!w || i < 0 || i >= w->boxes->count) return;
    BoxSystem* b = w->boxes; b->is_static[i] = st;
    if (st) b->velocities[i] = b->angular_velocities[i] = b->forces[i] = b->torques[i] = (Vec3){0, 0, 0};
}
#define WAKE(sys, i) do { if ((sys)->is_sleeping[i]) { (sys)->is_sleeping[i] = false; (sys)->sleep_timers[i] = 0.0f; } } while (0)
void AddSphereForce(PhysicsWorld* w, int i, Vec3 f) { if (!w || i < 0 || i >= w->spheres->count) return; WAKE(w->spheres, i); w->spheres->forces[i] = Vec3Add(w->spheres->forces[i], f); }
void AddSphereTorque(PhysicsWorld* w, int i, Vec3 t) { if (!w || i < 0 || i >= w->spheres->count) return; WAKE(w->spheres, i); w->spheres->torques[i] = Vec3Add(w->spheres->torques[i], t); }
void AddBoxForce(PhysicsWorld* w, int i, Vec3 f) { if (!w || i < 0 || i >= w->boxes->count) return; WAKE(w->boxes, i); w->boxes->forces[i] = Vec3Add(w->boxes->forces[i], f); }
void AddBoxTorque(PhysicsWorld* w, int i, Vec3 t) { if (!w || i < 0 || i >= w->boxes->count) return; WAKE(w->boxes, i); w->boxes->torques[i] = Vec3Add(w->boxes->torques[i], t); }
void SetSphereCollisionCallback(PhysicsWorld* w, int i, OnCollision cb) { if (w && i >= 0 && i < w->spheres->count) w->spheres->collision_callbacks[i] = cb; }
void SetBoxCollisionCallback(PhysicsWorld* w, int i, OnCollision cb) { if (w && i >= 0 && i < w->boxes->count) w->boxes->collision_callbacks[i] = cb; }

/* ---- joints: an EXTENSION of this repo (include/ballbox_ext.h: BallboxAddSpring / BallboxAddBallJoint /
 * BallboxAddDistanceJoint / BallboxAddHingeJoint); the reference only lists "constraints, like hinges, springs" as
 * planned (README.md:25).  This is the SEQUENTIAL definition the CUDA path is checked against bit for bit; nothing pins it
 * but itself.
 *   spring         : a force element, applied in creation order before gravity in ApplyForces
 *   ball joint     : three bilateral velocity rows (world x, y, z) that hold two anchor points together
 *   distance joint : one bilateral row along the line between the two anchor points that keeps their distance (a rod)
 *   hinge          : two ball joints on the hinge axis
 * Joint rows are presolved after the contacts and swept after the contacts in every solver iteration, joints in creation
 * order.  Anchors are stored in the body frame.  A sphere has no orientation in this engine (SphereSystem has no rotations),
 * so a sphere is always attached at its CENTRE (the anchor given for a sphere end is replaced by the centre).  Body type
 * -1 = the world (a fixed point). */
#define ORACLE_JOINT_WORLDS 512
#define ORACLE_JOINT_BALL 0
#define ORACLE_JOINT_ROD  1
typedef struct { int ta, ia, tb, ib; Vec3 la, lb; float rest, k, c; } OracleSpring;
typedef struct { int kind, ta, ia, tb, ib; Vec3 la, lb; float rest; Vec3 rA, rB, dir; float mass[3], bias[3], imp[3]; } OracleJoint;
static struct { PhysicsWorld* w; OracleSpring* sp; int nsp, csp; OracleJoint* jt; int njt, cjt; } g_joints[ORACLE_JOINT_WORLDS];
static int joint_slot(PhysicsWorld* w, int create)
{
    for (int i = 0; i < ORACLE_JOINT_WORLDS; i++) if (g_joints[i].w == w) return i;
    if (create) for (int i = 0; i < ORACLE_JOINT_WORLDS; i++) if (!g_joints[i].w) { g_joints[i].w = w; return i; }
    return -1;
}
static void joints_forget(PhysicsWorld* w)
{
    int i = joint_slot(w, 0);
    if (i >= 0) { free(g_joints[i].sp); free(g_joints[i].jt); memset(&g_joints[i], 0, sizeof(g_joints[i])); }
}
void BallboxClearJoints(PhysicsWorld* w) { joints_forget(w); }
static bool joint_body_ok(PhysicsWorld* w, int t, int i)
{
    if (t == -1) return true;
    if (t == 0) return i >= 0 && i < w->spheres->count;
    if (t == 1) return i >= 0 && i < w->boxes->count;
    return false;
}
/* world point -> body frame (boxes rotate; a sphere is attached at its centre; the world is a fixed point) */
static Vec3 joint_to_local(PhysicsWorld* w, int t, int i, Vec3 p)
{
    if (t == 0) return (Vec3){0, 0, 0};
    if (t == 1) return QuatRotateVec3(QuatConjugate(w->boxes->rotations[i]), Vec3Sub(p, w->boxes->positions[i]));
    return p;
}
static Vec3 joint_to_world(PhysicsWorld* w, int t, int i, Vec3 l)
{
    if (t == 0) return Vec3Add(w->spheres->positions[i], l);
    if (t == 1) return Vec3Add(w->boxes->positions[i], QuatRotateVec3(w->boxes->rotations[i], l));
    return l;
}
int BallboxAddSpring(PhysicsWorld* w, int ta, int ia, int tb, int ib, Vec3 anchorA, Vec3 anchorB, float rest, float k, float c)
{
    if (!w || ta == -1 || !joint_body_ok(w, ta, ia) || !joint_body_ok(w, tb, ib)) return -1;
    int g = joint_slot(w, 1); if (g < 0) return -1;
    if (g_joints[g].nsp == g_joints[g].csp) { g_joints[g].csp = g_joints[g].csp ? 2 * g_joints[g].csp : 16; g_joints[g].sp = (OracleSpring*)realloc(g_joints[g].sp, sizeof(OracleSpring) * (size_t)g_joints[g].csp); }
    OracleSpring* s = &g_joints[g].sp[g_joints[g].nsp];
    s->ta = ta; s->ia = ia; s->tb = tb; s->ib = tb == -1 ? -1 : ib; s->rest = rest; s->k = k; s->c = c;
    s->la = joint_to_local(w, ta, ia, anchorA); s->lb = joint_to_local(w, tb, ib, anchorB);
    return g_joints[g].nsp++;
}
static int joint_add(PhysicsWorld* w, int kind, int ta, int ia, int tb, int ib, Vec3 anchorA, Vec3 anchorB)
{
    if (!w || ta == -1 || !joint_body_ok(w, ta, ia) || !joint_body_ok(w, tb, ib)) return -1;
    int g = joint_slot(w, 1); if (g < 0) return -1;
    if (g_joints[g].njt == g_joints[g].cjt) { g_joints[g].cjt = g_joints[g].cjt ? 2 * g_joints[g].cjt : 16; g_joints[g].jt = (OracleJoint*)realloc(g_joints[g].jt, sizeof(OracleJoint) * (size_t)g_joints[g].cjt); }
    OracleJoint* j = &g_joints[g].jt[g_joints[g].njt];
    memset(j, 0, sizeof(*j));
    j->kind = kind; j->ta = ta; j->ia = ia; j->tb = tb; j->ib = tb == -1 ? -1 : ib;
    j->la = joint_to_local(w, ta, ia, anchorA); j->lb = joint_to_local(w, tb, ib, anchorB);
    j->rest = Vec3Length(Vec3Sub(joint_to_world(w, j->tb, j->ib, j->lb), joint_to_world(w, j->ta, j->ia, j->la)));   /* rods keep the distance they were created with */
    return g_joints[g].njt++;
}
int BallboxAddBallJoint(PhysicsWorld* w, int ta, int ia, int tb, int ib, Vec3 anchor) { return joint_add(w, ORACLE_JOINT_BALL, ta, ia, tb, ib, anchor, anchor); }
int BallboxAddDistanceJoint(PhysicsWorld* w, int ta, int ia, int tb, int ib, Vec3 anchorA, Vec3 anchorB) { return joint_add(w, ORACLE_JOINT_ROD, ta, ia, tb, ib, anchorA, anchorB); }
int BallboxAddHingeJoint(PhysicsWorld* w, int ta, int ia, int tb, int ib, Vec3 anchor, Vec3 axis, float half_span)
{
    Vec3 d = Vec3Scale(Vec3Normalize(axis), half_span);
    int first = BallboxAddBallJoint(w, ta, ia, tb, ib, Vec3Sub(anchor, d));
    if (first < 0) return -1;
    return BallboxAddBallJoint(w, ta, ia, tb, ib, Vec3Add(anchor, d)) < 0 ? -1 : first;
}
int BallboxSpringCount(PhysicsWorld* w) { int g = joint_slot(w, 0); return g < 0 ? 0 : g_joints[g].nsp; }
int BallboxJointCount(PhysicsWorld* w) { int g = joint_slot(w, 0); return g < 0 ? 0 : g_joints[g].njt; }

static void joint_body_state(PhysicsWorld* w, int t, int i, Vec3* pos, Vec3* v, Vec3* om, bool* dyn)
{
    *v = *om = (Vec3){0, 0, 0}; *dyn = false;
    if (t == 0) { *pos = w->spheres->positions[i]; if (!w->spheres->is_static[i]) { *dyn = true; *v = w->spheres->velocities[i]; *om = w->spheres->angular_velocities[i]; } }
    else if (t == 1) { *pos = w->boxes->positions[i]; if (!w->boxes->is_static[i]) { *dyn = true; *v = w->boxes->velocities[i]; *om = w->boxes->angular_velocities[i]; } }
}
static void spring_push(PhysicsWorld* w, int t, int i, Vec3 F, Vec3 T)
{
    if (t == 0) { SphereSystem* s = w->spheres; WAKE(s, i); s->forces[i] = Vec3Add(s->forces[i], F); s->torques[i] = Vec3Add(s->torques[i], T); }
    else { BoxSystem* b = w->boxes; WAKE(b, i); b->forces[i] = Vec3Add(b->forces[i], F); b->torques[i] = Vec3Add(b->torques[i], T); }
}
/* spring-dampers: forces and torques on both ends, in creation order, before gravity is added */
static void stage_springs(PhysicsWorld* w)
{
    int g = joint_slot(w, 0); if (g < 0) return;
    for (int n = 0; n < g_joints[g].nsp; n++) {
        const OracleSpring* s = &g_joints[g].sp[n];
        Vec3 posA, vA, oA, posB = s->lb, vB, oB; bool dynA, dynB;
        joint_body_state(w, s->ta, s->ia, &posA, &vA, &oA, &dynA);
        joint_body_state(w, s->tb, s->ib, &posB, &vB, &oB, &dynB);
        Vec3 pA = joint_to_world(w, s->ta, s->ia, s->la), pB = joint_to_world(w, s->tb, s->ib, s->lb);
        Vec3 d = Vec3Sub(pB, pA);
        float len = Vec3Length(d);
        if (len < 1e-6f) continue;
        Vec3 dir = Vec3Scale(d, 1.0f / len);
        Vec3 rA = Vec3Sub(pA, posA), rB = Vec3Sub(pB, posB);
        Vec3 velA = Vec3Add(vA, Vec3Cross(oA, rA)), velB = Vec3Add(vB, Vec3Cross(oB, rB));
        float f = s->k * (len - s->rest) + s->c * Vec3Dot(Vec3Sub(velB, velA), dir);
        Vec3 F = Vec3Scale(dir, f), Fm = Vec3Scale(F, -1.0f);
        if (dynA) spring_push(w, s->ta, s->ia, F, Vec3Cross(rA, F));
        if (dynB) spring_push(w, s->tb, s->ib, Fm, Vec3Cross(rB, Fm));
    }
}

/* ===== stage 1-2: ApplyForces :1820-1838, IntegrateVelocities :1841-1925 ==== */
static void stage_forces_velocities(PhysicsWorld* w)
{
    float dt = w->dt, ld = w->settings.linear_damping, ad = w->settings.angular_damping;
    SphereSystem* s = w->spheres; BoxSystem* b = w->boxes;
    for (int i = 0; i < s->count; i++) {
        if (!s->is_static[i]) {
            s->forces[i] = Vec3Add(s->forces[i], Vec3Scale(w->gravity, s->masses[i]));
            if (!s->is_sleeping[i]) {
                s->velocities[i] = Vec3Add(s->velocities[i], Vec3Scale(Vec3Scale(s->forces[i], 1.0f / s->masses[i]), dt));
                s->angular_velocities[i] = Vec3Add(s->angular_velocities[i], Vec3Scale(Vec3Scale(s->torques[i], 1.0f / s->inertias[i]), dt));
                s->velocities[i] = Vec3Scale(s->velocities[i], ld);
                s->angular_velocities[i] = Vec3Scale(s->angular_velocities[i], ad);
            }
        }
        s->forces[i] = s->torques[i] = (Vec3){0, 0, 0};
    }
    for (int i = 0; i < b->count; i++) {
        if (!b->is_static[i]) {
            b->forces[i] = Vec3Add(b->forces[i], Vec3Scale(w->gravity, b->masses[i]));
            if (!b->is_sleeping[i]) {
                b->velocities[i] = Vec3Add(b->velocities[i], Vec3Scale(Vec3Scale(b->forces[i], 1.0f / b->masses[i]), dt));
                Vec3 I = b->inertias[i], T = b->torques[i], al;
                al.x = (I.x > 1e-6f) ? T.x / I.x : 0.0f;                 /* world-axis, component-wise :1886-1891 */
                al.y = (I.y > 1e-6f) ? T.y / I.y : 0.0f;
                al.z = (I.z > 1e-6f) ? T.z / I.z : 0.0f;
                b->angular_velocities[i] = Vec3Add(b->angular_velocities[i], Vec3Scale(al, dt));
                b->velocities[i] = Vec3Scale(b->velocities[i], ld);
                b->angular_velocities[i] = Vec3Scale(b->angular_velocities[i], ad);
            }
        }
        b->forces[i] = b->torques[i] = (Vec3){0, 0, 0};
    }
}

/* ===== narrowphase ======================================================== */
typedef struct { Vec3 c, h, ax[3]; } Obb;
static float comp(Vec3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
static Obb obb_of(Vec3 p, Vec3 h, Quat q)                                    /* CreateOBB :826-834 */
{
    Obb o; o.c = p; o.h = h;
    o.ax[0] = QuatRotateVec3(q, (Vec3){1, 0, 0}); o.ax[1] = QuatRotateVec3(q, (Vec3){0, 1, 0}); o.ax[2] = QuatRotateVec3(q, (Vec3){0, 0, 1});
    return o;
}
static void obb_corners(const Obb* o, Vec3 v[8])                             /* GetOBBVertices :837-850, same sums */
{
    Vec3 X = Vec3Scale(o->ax[0], o->h.x), Y = Vec3Scale(o->ax[1], o->h.y), Z = Vec3Scale(o->ax[2], o->h.z);
    v[0] = Vec3Add(Vec3Add(Vec3Add(o->c, X), Y), Z);
    v[1] = Vec3Add(Vec3Add(Vec3Sub(o->c, Z), X), Y);
    v[2] = Vec3Add(Vec3Add(Vec3Add(o->c, X), Z), Vec3Scale(Y, -1));
    v[3] = Vec3Add(Vec3Sub(Vec3Add(o->c, X), Y), Vec3Scale(Z, -1));
    v[4] = Vec3Add(Vec3Add(Vec3Add(o->c, Y), Z), Vec3Scale(X, -1));
    v[5] = Vec3Add(Vec3Sub(Vec3Add(o->c, Y), Z), Vec3Scale(X, -1));
    v[6] = Vec3Add(Vec3Sub(Vec3Add(o->c, Z), X), Vec3Scale(Y, -1));
    v[7] = Vec3Sub(Vec3Sub(Vec3Sub(o->c, X), Y), Z);
}
static void obb_interval(const Obb* o, Vec3 a, float* lo, float* hi)        /* ProjectOBB :853-861 */
{
    float c = Vec3Dot(o->c, a);
    float e = o->h.x * fabsf(Vec3Dot(o->ax[0], a)) + o->h.y * fabsf(Vec3Dot(o->ax[1], a)) + o->h.z * fabsf(Vec3Dot(o->ax[2], a));
    *lo = c - e; *hi = c + e;
}
static bool axis_overlap(const Obb* a, const Obb* b, Vec3 axis, float* depth) /* OverlapOnAxis :864-884 */
{
    float l1, h1, l2, h2;
    obb_interval(a, axis, &l1, &h1); obb_interval(b, axis, &l2, &h2);
    if (h1 < l2 || h2 < l1) return false;
    *depth = fminf(h1, h2) - fmaxf(l1, l2);
    if (l1 < l2) *depth *= -1.0f;
    return true;
}
static Vec3 deepest_vertex_on_face(const Obb* vo, const Obb* fo, float* best)   /* VertexFaceCollision :908-950 */
{
    Vec3 v[8]; obb_corners(vo, v);
    Vec3 pick = {0, 0, 0}; *best = 999999.0f;
    for (int i = 0; i < 3; i++) {
        Vec3 n = fo->ax[i], U = Vec3Normalize(fo->ax[(i + 1) % 3]), V = Vec3Normalize(fo->ax[(i + 2) % 3]);
        float uh = comp(fo->h, (i + 1) % 3), vh = comp(fo->h, (i + 2) % 3);
        for (int d = -1; d <= 1; d += 2) {
            Vec3 dn = Vec3Scale(n, (float)d);
            Vec3 fc = Vec3Add(fo->c, Vec3Scale(n, comp(fo->h, i) * (float)d));
            for (int j = 0; j < 8; j++) {
                float sd = Vec3Dot(Vec3Sub(v[j], fc), dn);
                if (sd < 0.0f) {
                    Vec3 rel = Vec3Sub(v[j], fc);
                    if (fabsf(Vec3Dot(rel, U)) <= uh && fabsf(Vec3Dot(rel, V)) <= vh && -sd < *best) { *best = -sd; pick = v[j]; }
                }
            }
        }
    }
    return pick;
}
static float seg_seg_dist2(Vec3 p1, Vec3 q1, Vec3 p2, Vec3 q2)               /* SquaredDistanceBetweenEdges :953-990 */
{
    Vec3 d1 = Vec3Sub(q1, p1), d2 = Vec3Sub(q2, p2), r = Vec3Sub(p1, p2);
    float a = Vec3Dot(d1, d1), e = Vec3Dot(d2, d2), f = Vec3Dot(d2, r), s, t;
    if (a <= 1e-6f && e <= 1e-6f) return Vec3Dot(r, r);
    if (a <= 1e-6f) { s = 0.0f; t = fmaxf(0.0f, fminf(1.0f, f / e)); }
    else {
        float c = Vec3Dot(d1, r);
        if (e <= 1e-6f) { t = 0.0f; s = fmaxf(0.0f, fminf(1.0f, -c / a)); }
        else {
            float b = Vec3Dot(d1, d2), den = a * e - b * b;
            s = (den != 0.0f) ? fmaxf(0.0f, fminf(1.0f, (b * f - c * e) / den)) : 0.0f;
            t = fmaxf(0.0f, fminf(1.0f, (b * s + f) / e));
        }
    }
    Vec3 df = Vec3Sub(Vec3Add(p1, Vec3Scale(d1, s)), Vec3Add(p2, Vec3Scale(d2, t)));
    return Vec3Dot(df, df);
}
static Vec3 edge_case_point(const Obb* a, const Obb* b)                       /* ClosestPointBetweenEdges :993-1034 */
{
    Vec3 va[8], vb[8]; obb_corners(a, va); obb_corners(b, vb);
    int pick[3] = {1, 2, 4};
    float best = 999999.0f; Vec3 out = {0, 0, 0};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        float d = seg_seg_dist2(va[0], Vec3Add(va[0], Vec3Sub(va[pick[i]], va[0])), vb[0], Vec3Add(vb[0], Vec3Sub(vb[pick[j]], vb[0])));
        if (d < best) { best = d; out = Vec3Scale(Vec3Add(a->c, b->c), 0.5f); }   /* always the centre midpoint (:1028) */
    }
    return out;
}
/* SATCollision :1040-1111 + normal flip of CheckBoxCollisionWithData :1121-1129 */
static bool sat(const Obb* A, const Obb* B, Vec3* pt, Vec3* nrm, float* pen)
{
    Vec3 axes[15]; int n = 0;
    for (int i = 0; i < 3; i++) axes[n++] = A->ax[i];
    for (int i = 0; i < 3; i++) axes[n++] = B->ax[i];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        Vec3 c = Vec3Cross(A->ax[i], B->ax[j]); float l = Vec3Length(c);
        if (l > 1e-6f) axes[n++] = Vec3Scale(c, 1.0f / l);
    }
    float best = 999999.0f; Vec3 bax = {0, 0, 0}; int which = -1;
    for (int i = 0; i < n; i++) {
        float d = 0.0f;
        if (!axis_overlap(A, B, axes[i], &d)) return false;
        if (fabsf(d) < fabsf(best)) { best = d; bax = Vec3Scale(axes[i], (d < 0) ? -1.0f : 1.0f); which = i; }
    }
    *pen = fabsf(best); *nrm = bax;
    if (which >= 0 && which <= 5) {
        float d1, d2; Vec3 p1 = deepest_vertex_on_face(A, B, &d1), p2 = deepest_vertex_on_face(B, A, &d2);
        *pt = (d1 < d2) ? p1 : p2;
    } else if (which >= 6) *pt = edge_case_point(A, B);
    else *pt = Vec3Scale(Vec3Add(A->c, B->c), 0.5f);
    if (Vec3Dot(*nrm, Vec3Sub(B->c, A->c)) < 0.0f) *nrm = Vec3Scale(*nrm, -1.0f);
    return true;
}
static bool in_grown_box(Vec3 p, Vec3 c, Vec3 h, Quat q, float tol, float* slack)   /* :2598-2613 */
{
    Vec3 l = QuatRotateVec3(QuatConjugate(q), Vec3Sub(p, c));
    if (fabsf(l.x) <= h.x + tol && fabsf(l.y) <= h.y + tol && fabsf(l.z) <= h.z + tol) {
        float px = h.x - fabsf(l.x), py = h.y - fabsf(l.y), pz = h.z - fabsf(l.z);
        *slack = fminf(px, fminf(py, pz));
        return true;
    }
    return false;
}
/* GenerateBoxBoxManifold :2550-2724 + SelectOptimalContactPoints :2726-2843 */
static int box_box(PhysicsWorld* w, int ia, int ib, Vec3 pts[8], float pens[8], Vec3* normal)
{
    BoxSystem* b = w->boxes;
    Vec3 pa = b->positions[ia], ha = b->sizes[ia], pb = b->positions[ib], hb = b->sizes[ib];
    Quat qa = b->rotations[ia], qb = b->rotations[ib];
    Obb A = obb_of(pa, ha, qa), B = obb_of(pb, hb, qb);
    Vec3 cp[24]; float cd[24]; int n = 0;
    Vec3 sp; float sd;
    if (!sat(&A, &B, &sp, normal, &sd)) return 0;
    cp[0] = sp; cd[0] = sd; n = 1;
    float tol = w->settings.manifold_contact_tolerance, ptol = w->settings.manifold_penetration_tolerance, sl;
    Vec3 va[8], vb[8]; obb_corners(&A, va); obb_corners(&B, vb);
    for (int i = 0; i < 8 && n < 24; i++) if (in_grown_box(va[i], pb, hb, qb, tol, &sl) && sl > ptol) { cp[n] = va[i]; cd[n] = sl; n++; }
    for (int i = 0; i < 8 && n < 24; i++) if (in_grown_box(vb[i], pa, ha, qa, tol, &sl) && sl > ptol) { cp[n] = vb[i]; cd[n] = sl; n++; }
    if (n < 3) {                                                               /* literal edge table :2671-2680 */
        for (int e = 0; e < 12 && n < 24; e++) {
            Vec3 s, t;
            if (e < 4) { s = va[e]; t = va[(e + 1) % 4]; }
            else if (e < 8) { s = va[e]; t = va[4 + ((e - 4 + 1) % 4)]; }
            else { s = va[e - 8]; t = va[e - 8 + 4]; }
            Vec3 mid = Vec3Scale(Vec3Add(s, t), 0.5f);
            if (in_grown_box(mid, pb, hb, qb, tol, &sl) && sl > ptol) { cp[n] = mid; cd[n] = sl; n++; }
        }
    }
    int maxc = w->settings.manifold_max_contacts;
    if (maxc < 1) maxc = 1;
    if (maxc > MAX_MANIFOLD_CONTACTS) maxc = MAX_MANIFOLD_CONTACTS;
    if (n <= 2) { for (int i = 0; i < n; i++) { pts[i] = cp[i]; pens[i] = cd[i]; } return n; }
    int deep = 0;
    for (int i = 1; i < n; i++) if (cd[i] > cd[deep]) deep = i;
    pts[0] = cp[deep]; pens[0] = cd[deep];
    if (n <= maxc) { int k = 1; for (int i = 0; i < n; i++) if (i != deep) { pts[k] = cp[i]; pens[k] = cd[i]; k++; } return n; }
    Vec3 ctr = {0, 0, 0};
    for (int i = 0; i < n; i++) ctr = Vec3Add(ctr, cp[i]);
    ctr = Vec3Scale(ctr, 1.0f / n);
    int rem = maxc - 1; float bd[8]; int bi[8];
    for (int j = 0; j < rem; j++) { bd[j] = -1.0f; bi[j] = -1; }
    for (int i = 0; i < n; i++) {
        if (i == deep) continue;
        float d = Vec3Length(Vec3Sub(cp[i], ctr));
        for (int j = 0; j < rem; j++) if (d > bd[j]) {
            for (int k = rem - 1; k > j; k--) { bd[k] = bd[k - 1]; bi[k] = bi[k - 1]; }
            bd[j] = d; bi[j] = i; break;
        }
    }
    int np = 1;
    for (int j = 0; j < rem; j++) if (bi[j] >= 0) { pts[np] = cp[bi[j]]; pens[np] = cd[bi[j]]; np++; }
    return np;
}
/* CheckSphereCollisionWithData :451-471 */
static bool sphere_sphere(Vec3 pa, float ra, Vec3 pb, float rb, CollisionContact* c)
{
    Vec3 n = Vec3Sub(pb, pa); float d = Vec3Length(n), rs = ra + rb;
    if (!(d < rs)) return false;
    c->contact_normal = (d > 1e-6f) ? Vec3Scale(n, 1.0f / d) : (Vec3){1.0f, 0.0f, 0.0f};
    c->contact_point = Vec3Add(pa, Vec3Scale(c->contact_normal, ra));
    c->penetration = rs - d;
    return true;
}
/* CheckSphereBoxCollisionWithData :1565-1626 */
static bool sphere_box(Vec3 sp, float sr, Vec3 bp, Vec3 bh, Quat bq, CollisionContact* c)
{
    Vec3 loc = QuatRotateVec3((Quat){ -bq.x, -bq.y, -bq.z, bq.w }, Vec3Sub(sp, bp));
    Vec3 cl = { fmaxf(-bh.x, fminf(loc.x, bh.x)), fmaxf(-bh.y, fminf(loc.y, bh.y)), fmaxf(-bh.z, fminf(loc.z, bh.z)) };
    Vec3 df = Vec3Sub(loc, cl); float d2 = Vec3Dot(df, df);
    if (!(d2 <= sr * sr)) return false;
    float d = sqrtf(d2); Vec3 n;
    if (d > 1e-6f) n = Vec3Scale(df, 1.0f / d);
    else {
        float ax = fabsf(loc.x), ay = fabsf(loc.y), az = fabsf(loc.z);
        if (ax >= ay && ax >= az) n = loc.x > 0 ? (Vec3){1, 0, 0} : (Vec3){-1, 0, 0};
        else if (ay >= az) n = loc.y > 0 ? (Vec3){0, 1, 0} : (Vec3){0, -1, 0};
        else n = loc.z > 0 ? (Vec3){0, 0, 1} : (Vec3){0, 0, -1};
    }
    n = QuatRotateVec3(bq, n);
    if (Vec3Dot(n, Vec3Sub(bp, sp)) < 0.0f) n = Vec3Scale(n, -1.0f);
    c->contact_normal = n;
    c->contact_point = Vec3Add(sp, Vec3Scale(n, -sr));
    c->penetration = sr - d;
    return true;
}
static Vec3 sdf_gradient(PhysicsWorld* w, Vec3 p)                             /* SDF_EstimateNormal :2318-2331 */
{
    float e = w->settings.sdf_normal_epsilon;
    Vec3 ex = {e, 0.0f, 0.0f}, ey = {0.0f, e, 0.0f}, ez = {0.0f, 0.0f, e};
    float gx = (w->SDFCollider(Vec3Add(p, ex)) - w->SDFCollider(Vec3Sub(p, ex))) / (2.0f * e);
    float gy = (w->SDFCollider(Vec3Add(p, ey)) - w->SDFCollider(Vec3Sub(p, ey))) / (2.0f * e);
    float gz = (w->SDFCollider(Vec3Add(p, ez)) - w->SDFCollider(Vec3Sub(p, ez))) / (2.0f * e);
    return Vec3Normalize((Vec3){gx, gy, gz});
}

/* ===== stage 3: CollectCollisions :1979-2115 ============================== */
static CollisionContact* next_slot(PhysicsWorld* w) { return &w->collisions->contacts[w->collisions->count]; }
static bool full(PhysicsWorld* w) { return w->collisions->count >= w->collisions->capacity; }

static float oracle_box_bound(const BoxSystem* b, int i)
{
    Quat q = b->rotations[i]; float q2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (q2 < 1.0f) q2 = 1.0f;
    return Vec3Length(b->sizes[i]) * q2 * 1.0005f + 1e-5f;
}

static void stage_detect(PhysicsWorld* w, int pruned)
{
    SphereSystem* S = w->spheres; BoxSystem* B = w->boxes; CollisionCollection* C = w->collisions;
    C->count = 0;
    PruneGrid *gs = NULL, *gb = NULL; float* brad = NULL; int* cand = NULL; int cap = 0;
    if (pruned) {
        brad = (float*)malloc(sizeof(float) * (size_t)(B->count > 0 ? B->count : 1));
        float cut = 0.0f; int any = 0;
        for (int i = 0; i < B->count; i++) brad[i] = oracle_box_bound(B, i);
        for (int i = 0; i < S->count; i++) if (!S->is_static[i]) { any = 1; if (S->radii[i] > cut) cut = S->radii[i]; }
        for (int i = 0; i < B->count; i++) if (!B->is_static[i]) { any = 1; if (brad[i] > cut) cut = brad[i]; }
        if (!any) { for (int i = 0; i < S->count; i++) if (S->radii[i] > cut) cut = S->radii[i];
                    for (int i = 0; i < B->count; i++) if (brad[i] > cut) cut = brad[i]; }
        if (!(cut > 0.0f)) cut = 1.0f;
        gs = pg_build(S->count, (const float*)S->positions, S->radii, cut);
        gb = pg_build(B->count, (const float*)B->positions, brad, cut);
    }
    /* phase 0: sphere-sphere :1986-2008 */
    for (int i = 0; i < S->count; i++) {
        int nc = pruned ? pg_query(gs, (const float*)&S->positions[i], S->radii[i], i + 1, &cand, &cap) : S->count - i - 1;
        for (int t = 0; t < nc; t++) {
            int j = pruned ? cand[t] : i + 1 + t;
            if (full(w)) break;
            CollisionContact* c = next_slot(w);
            if (sphere_sphere(S->positions[i], S->radii[i], S->positions[j], S->radii[j], c)) {
                c->bodyA_type = 0; c->bodyA_index = i; c->bodyB_type = 0; c->bodyB_index = j; C->count++;
                if (S->collision_callbacks[i]) S->collision_callbacks[i](i, 0, j, c);
                if (S->collision_callbacks[j]) S->collision_callbacks[j](j, 0, i, c);
            }
        }
    }
    /* phase 1: box-box manifolds :2011-2047 */
    for (int i = 0; i < B->count; i++) {
        int nc = pruned ? pg_query(gb, (const float*)&B->positions[i], brad[i], i + 1, &cand, &cap) : B->count - i - 1;
        for (int t = 0; t < nc; t++) {
            int j = pruned ? cand[t] : i + 1 + t;
            Vec3 pts[8], nrm; float pens[8];
            int np = box_box(w, i, j, pts, pens, &nrm);
            if (np > 0) {
                int first = C->count;
                for (int k = 0; k < np; k++) {
                    if (full(w)) break;
                    CollisionContact* c = next_slot(w);
                    c->bodyA_type = 1; c->bodyA_index = i; c->bodyB_type = 1; c->bodyB_index = j;
                    c->contact_point = pts[k]; c->contact_normal = nrm; c->penetration = pens[k]; C->count++;
                }
                (void)first;
                if (B->collision_callbacks[i]) B->collision_callbacks[i](i, 1, j, &C->contacts[C->count - np]);
                if (B->collision_callbacks[j]) B->collision_callbacks[j](j, 1, i, &C->contacts[C->count - np]);
            }
        }
    }
    /* phase 2: sphere-box :2050-2072 */
    for (int i = 0; i < S->count; i++) {
        int nc = pruned ? pg_query(gb, (const float*)&S->positions[i], S->radii[i], 0, &cand, &cap) : B->count;
        for (int t = 0; t < nc; t++) {
            int j = pruned ? cand[t] : t;
            if (full(w)) break;
            CollisionContact* c = next_slot(w);
            if (sphere_box(S->positions[i], S->radii[i], B->positions[j], B->sizes[j], B->rotations[j], c)) {
                c->bodyA_type = 0; c->bodyA_index = i; c->bodyB_type = 1; c->bodyB_index = j; C->count++;
                if (S->collision_callbacks[i]) S->collision_callbacks[i](i, 1, j, c);
                if (B->collision_callbacks[j]) B->collision_callbacks[j](j, 0, i, c);
            }
        }
    }
    if (pruned) { free(cand); pg_free(gs); pg_free(gb); free(brad); }
    if (!(w->usesSDF && w->SDFCollider)) return;
    /* phase 3: sphere-SDF :2074-2106 */
    for (int i = 0; i < S->count; i++) {
        if (full(w)) break;
        Vec3 ctr = S->positions[i]; float r = S->radii[i], d = w->SDFCollider(ctr);
        if (d < r) {
            CollisionContact* c = next_slot(w);
            c->penetration = r - d;
            c->contact_normal = Vec3Scale(sdf_gradient(w, ctr), -1.0f);
            c->contact_point = Vec3Sub(ctr, Vec3Scale(c->contact_normal, r));
            c->bodyA_type = 0; c->bodyA_index = i; c->bodyB_type = 2; c->bodyB_index = 0; C->count++;
            if (S->collision_callbacks[i]) S->collision_callbacks[i](i, 2, 0, c);
        }
    }
    /* phase 4: box-SDF, 8 corners then 12 "edge" midpoints :2413-2477 */
    static const int ep[12][2] = { {0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7} };
    for (int i = 0; i < B->count; i++) {
        Obb o = obb_of(B->positions[i], B->sizes[i], B->rotations[i]);
        Vec3 v[8]; obb_corners(&o, v);
        for (int k = 0; k < 20; k++) {
            if (k == 8 && full(w)) break;                                      /* second loop's own capacity check */
            if (full(w)) break;
            Vec3 p = (k < 8) ? v[k] : Vec3Scale(Vec3Add(v[ep[k - 8][0]], v[ep[k - 8][1]]), 0.5f);
            float d = w->SDFCollider(p);
            if (d < -0.001f) {
                CollisionContact* c = next_slot(w);
                c->penetration = -d;
                c->contact_normal = Vec3Scale(sdf_gradient(w, p), -1.0f);
                c->contact_point = p;
                c->bodyA_type = 1; c->bodyA_index = i; c->bodyB_type = 2; c->bodyB_index = 0; C->count++;
            }
        }
    }
}

/* ===== stage 4: SolveConstraints :1176-1226, :1295-1489 =================== */
typedef struct { Vec3 pos; float im; Vec3 iI; bool dyn; int type, idx; } BodyRef;
static BodyRef body_ref(PhysicsWorld* w, int type, int idx, Vec3 contact_point)
{
    BodyRef r; r.type = type; r.idx = idx; r.dyn = false; r.im = 0.0f; r.iI = (Vec3){0, 0, 0}; r.pos = contact_point;
    if (type == 0) {
        r.pos = w->spheres->positions[idx];
        if (!w->spheres->is_static[idx]) { r.dyn = true; r.im = 1.0f / w->spheres->masses[idx]; float ii = 1.0f / w->spheres->inertias[idx]; r.iI = (Vec3){ii, ii, ii}; }
    } else if (type == 1) {
        r.pos = w->boxes->positions[idx];
        if (!w->boxes->is_static[idx]) { r.dyn = true; r.im = 1.0f / w->boxes->masses[idx]; Vec3 I = w->boxes->inertias[idx]; r.iI = (Vec3){1.0f / I.x, 1.0f / I.y, 1.0f / I.z}; }
    }
    return r;
}
static float eff_mass(const BodyRef* a, const BodyRef* b, Vec3 rA, Vec3 rB, Vec3 d)         /* :1321-1326 */
{
    Vec3 ra = Vec3Cross(rA, d), rb = Vec3Cross(rB, d);
    float s = a->im + b->im + Vec3Dot(ra, (Vec3){ra.x * a->iI.x, ra.y * a->iI.y, ra.z * a->iI.z})
                            + Vec3Dot(rb, (Vec3){rb.x * b->iI.x, rb.y * b->iI.y, rb.z * b->iI.z});
    return (s > 0.0f) ? 1.0f / s : 0.0f;
}
static Vec3 point_velocity(PhysicsWorld* w, const BodyRef* b, Vec3 p)                        /* :1233-1261 */
{
    if (!b->dyn) return (Vec3){0, 0, 0};
    Vec3 v = b->type == 0 ? w->spheres->velocities[b->idx] : w->boxes->velocities[b->idx];
    Vec3 om = b->type == 0 ? w->spheres->angular_velocities[b->idx] : w->boxes->angular_velocities[b->idx];
    return Vec3Add(v, Vec3Cross(om, Vec3Sub(p, b->pos)));
}
static Vec3 body_v(PhysicsWorld* w, const BodyRef* b) { return b->type == 0 ? w->spheres->velocities[b->idx] : w->boxes->velocities[b->idx]; }
static Vec3 body_om(PhysicsWorld* w, const BodyRef* b) { return b->type == 0 ? w->spheres->angular_velocities[b->idx] : w->boxes->angular_velocities[b->idx]; }
static void push(PhysicsWorld* w, const BodyRef* b, Vec3 J, Vec3 L)                          /* ApplyConstraintImpulse :1359-1408 */
{
    if (!b->dyn) return;
    if (b->type == 0) {
        SphereSystem* s = w->spheres; int i = b->idx;
        s->velocities[i] = Vec3Add(s->velocities[i], Vec3Scale(J, b->im));
        s->angular_velocities[i] = Vec3Add(s->angular_velocities[i], Vec3Scale(L, b->iI.x));
        if (s->is_sleeping[i]) { s->is_sleeping[i] = false; s->sleep_timers[i] = 0.0f; }
    } else {
        BoxSystem* x = w->boxes; int i = b->idx; Quat q = x->rotations[i];
        x->velocities[i] = Vec3Add(x->velocities[i], Vec3Scale(J, b->im));
        Vec3 l = QuatRotateVec3(QuatConjugate(q), L);
        Vec3 dl = { l.x * b->iI.x, l.y * b->iI.y, l.z * b->iI.z };
        x->angular_velocities[i] = Vec3Add(x->angular_velocities[i], QuatRotateVec3(q, dl));
        if (x->is_sleeping[i]) { x->is_sleeping[i] = false; x->sleep_timers[i] = 0.0f; }
    }
}
/* ---- warm starting: an EXTENSION of this repo (include/ballbox_ext.h, BallboxSetWarmStarting), not reference
 * behaviour; the reference zeroes the accumulated impulses (:1203-1206).  This is the sequential definition the CUDA
 * path is checked against; nothing pins it but itself. */
#define ORACLE_WARM_WORLDS 64
static struct { PhysicsWorld* w; float factor; ContactConstraint* prev; int n, cap; } g_warm[ORACLE_WARM_WORLDS];
static int warm_slot(PhysicsWorld* w, int create)
{
    for (int i = 0; i < ORACLE_WARM_WORLDS; i++) if (g_warm[i].w == w) return i;
    if (create) for (int i = 0; i < ORACLE_WARM_WORLDS; i++) if (!g_warm[i].w) { g_warm[i].w = w; return i; }
    return -1;
}
void OracleSetWarmStarting(PhysicsWorld* w, float factor)
{
    int i = warm_slot(w, 1);
    if (i >= 0) g_warm[i].factor = factor > 0.0f ? factor : 0.0f;
}
static void warm_forget(PhysicsWorld* w)
{
    int i = warm_slot(w, 0);
    if (i >= 0) { free(g_warm[i].prev); memset(&g_warm[i], 0, sizeof(g_warm[i])); }
}
static bool same_pair_ids(const ContactConstraint* a, const ContactConstraint* b)
{
    return a->bodyA_type == b->bodyA_type && a->bodyA_index == b->bodyA_index && a->bodyB_type == b->bodyB_type && a->bodyB_index == b->bodyB_index;
}
/* index k of constraint i among the consecutive records of its body pair */
static int run_index(const ContactConstraint* K, int i) { int k = 0; while (i - k - 1 >= 0 && same_pair_ids(&K[i], &K[i - k - 1])) k++; return k; }

static void stage_solve(PhysicsWorld* w)
{
    ConstraintCollection* K = w->constraints; CollisionCollection* C = w->collisions;
    const int ws = warm_slot(w, 0);
    const float warm = ws >= 0 ? g_warm[ws].factor : 0.0f;
    bool* matched = NULL;
    if (warm > 0.0f) {                                   /* keep the previous list: K is about to be overwritten */
        if (g_warm[ws].cap < K->capacity) { g_warm[ws].prev = (ContactConstraint*)realloc(g_warm[ws].prev, sizeof(ContactConstraint) * (size_t)K->capacity); g_warm[ws].cap = K->capacity; }
        memcpy(g_warm[ws].prev, K->constraints, sizeof(ContactConstraint) * (size_t)K->count);
        g_warm[ws].n = K->count;
        matched = (bool*)calloc((size_t)(C->count > 0 ? C->count : 1), sizeof(bool));
    }
    K->count = 0;
    for (int i = 0; i < C->count && K->count < K->capacity; i++) {                           /* :1184-1210 */
        CollisionContact* s = &C->contacts[i]; ContactConstraint* c = &K->constraints[K->count++];
        c->bodyA_type = s->bodyA_type; c->bodyA_index = s->bodyA_index; c->bodyB_type = s->bodyB_type; c->bodyB_index = s->bodyB_index;
        c->contact_point = s->contact_point; c->contact_normal = s->contact_normal; c->penetration = s->penetration;
        c->restitution = w->settings.restitution; c->friction = w->settings.friction;
        c->normal_impulse = 0.0f; c->tangent_impulse[0] = 0.0f; c->tangent_impulse[1] = 0.0f;
    }
    if (warm > 0.0f) {
        /* match by (body A, body B, k); both lists are in emission order but a plain search per pair start is enough here */
        const ContactConstraint* P = g_warm[ws].prev; const int np = g_warm[ws].n;
        int from = 0;
        for (int i = 0; i < K->count; i++) {
            ContactConstraint* c = &K->constraints[i];
            const int k = run_index(K->constraints, i);
            int j = -1;
            for (int t = 0; t < np; t++) {                                   /* resume where the last match was: O(n) in practice */
                int q = from + t; if (q >= np) q -= np;
                if (same_pair_ids(c, &P[q]) && run_index(P, q) == k) { j = q; break; }
            }
            if (j < 0) continue;
            from = j;
            c->normal_impulse = warm * P[j].normal_impulse;
            c->tangent_impulse[0] = warm * P[j].tangent_impulse[0];
            c->tangent_impulse[1] = warm * P[j].tangent_impulse[1];
            matched[i] = true;
        }
    }
    for (int i = 0; i < K->count; i++) {                                                     /* presolve :1300-1355 */
        ContactConstraint* c = &K->constraints[i];
        BodyRef a = body_ref(w, c->bodyA_type, c->bodyA_index, c->contact_point), b = body_ref(w, c->bodyB_type, c->bodyB_index, c->contact_point);
        Vec3 rA = Vec3Sub(c->contact_point, a.pos), rB = Vec3Sub(c->contact_point, b.pos), n = c->contact_normal;
        c->normal_mass = eff_mass(&a, &b, rA, rB, n);
        Vec3 t1 = (fabsf(n.x) >= 0.57735f) ? (Vec3){n.y, -n.x, 0.0f} : (Vec3){0.0f, n.z, -n.y};
        t1 = Vec3Normalize(t1);
        c->tangent[0] = t1; c->tangent[1] = Vec3Cross(n, t1);
        for (int j = 0; j < 2; j++) c->tangent_mass[j] = eff_mass(&a, &b, rA, rB, c->tangent[j]);
        c->position_bias = (w->settings.baumgarte_bias / w->dt) * fmaxf(0.0f, c->penetration - w->settings.allowed_penetration);
    }
    const int jg = joint_slot(w, 0);
    const int njt = jg >= 0 ? g_joints[jg].njt : 0;
    for (int n = 0; n < njt; n++) {                                                          /* joint rows: presolve */
        OracleJoint* j = &g_joints[jg].jt[n];
        Vec3 pA = joint_to_world(w, j->ta, j->ia, j->la), pB = joint_to_world(w, j->tb, j->ib, j->lb);
        BodyRef a = body_ref(w, j->ta, j->ia, pA), b = body_ref(w, j->tb == -1 ? 2 : j->tb, j->ib, pB);
        j->rA = Vec3Sub(pA, a.pos); j->rB = Vec3Sub(pB, b.pos);
        Vec3 C = Vec3Sub(pB, pA);
        if (j->kind == ORACLE_JOINT_ROD) {
            j->dir = Vec3Normalize(C);
            j->mass[0] = eff_mass(&a, &b, j->rA, j->rB, j->dir);
            j->bias[0] = -((w->settings.baumgarte_bias / w->dt) * (Vec3Length(C) - j->rest));
            j->imp[0] = 0.0f;
            continue;
        }
        for (int r = 0; r < 3; r++) {
            Vec3 e = { r == 0 ? 1.0f : 0.0f, r == 1 ? 1.0f : 0.0f, r == 2 ? 1.0f : 0.0f };
            j->mass[r] = eff_mass(&a, &b, j->rA, j->rB, e);
            j->bias[r] = -((w->settings.baumgarte_bias / w->dt) * comp(C, r));
            j->imp[r] = 0.0f;
        }
    }
    if (warm > 0.0f && w->settings.solver_iterations > 0) {
        for (int i = 0; i < K->count; i++) {
            if (!matched[i]) continue;
            ContactConstraint* c = &K->constraints[i];
            BodyRef a = body_ref(w, c->bodyA_type, c->bodyA_index, c->contact_point), b = body_ref(w, c->bodyB_type, c->bodyB_index, c->contact_point);
            Vec3 rA = Vec3Sub(c->contact_point, a.pos), rB = Vec3Sub(c->contact_point, b.pos);
            Vec3 P = Vec3Scale(c->contact_normal, c->normal_impulse), Pm = Vec3Scale(P, -1.0f);
            push(w, &a, Pm, Vec3Cross(rA, Pm));
            push(w, &b, P, Vec3Cross(rB, P));
            if (c->friction > 0.0f)
                for (int j = 0; j < 2; j++) {
                    Vec3 T = Vec3Scale(c->tangent[j], c->tangent_impulse[j]), Tm = Vec3Scale(T, -1.0f);
                    push(w, &a, Tm, Vec3Cross(rA, Tm));
                    push(w, &b, T, Vec3Cross(rB, T));
                }
        }
    }
    free(matched);
    for (int it = 0; it < w->settings.solver_iterations; it++) {                             /* :1223, :1414-1488 */
        for (int i = 0; i < K->count; i++) {
            ContactConstraint* c = &K->constraints[i];
            BodyRef a = body_ref(w, c->bodyA_type, c->bodyA_index, c->contact_point), b = body_ref(w, c->bodyB_type, c->bodyB_index, c->contact_point);
            Vec3 rA = Vec3Sub(c->contact_point, a.pos), rB = Vec3Sub(c->contact_point, b.pos);
            Vec3 rel = Vec3Sub(point_velocity(w, &b, c->contact_point), point_velocity(w, &a, c->contact_point));
            float vn = Vec3Dot(rel, c->contact_normal), vb = 0.0f;
            if (vn < -w->settings.velocity_threshold) vb = -c->restitution * vn;
            float jn = c->normal_mass * (-(vn + vb) + c->position_bias), old = c->normal_impulse;
            c->normal_impulse = fmaxf(0.0f, c->normal_impulse + jn);
            jn = c->normal_impulse - old;
            Vec3 P = Vec3Scale(c->contact_normal, jn), Pm = Vec3Scale(P, -1.0f);
            push(w, &a, Pm, Vec3Cross(rA, Pm));
            push(w, &b, P, Vec3Cross(rB, P));
            if (c->friction > 0.0f) {
                rel = Vec3Sub(point_velocity(w, &b, c->contact_point), point_velocity(w, &a, c->contact_point));
                for (int j = 0; j < 2; j++) {
                    float jt = c->tangent_mass[j] * (-Vec3Dot(rel, c->tangent[j]));
                    float lim = c->friction * c->normal_impulse, o = c->tangent_impulse[j];
                    c->tangent_impulse[j] = fmaxf(-lim, fminf(lim, c->tangent_impulse[j] + jt));
                    jt = c->tangent_impulse[j] - o;
                    Vec3 T = Vec3Scale(c->tangent[j], jt), Tm = Vec3Scale(T, -1.0f);
                    push(w, &a, Tm, Vec3Cross(rA, Tm));
                    push(w, &b, T, Vec3Cross(rB, T));
                }
            }
        }
        for (int n = 0; n < njt; n++) {                                                      /* joint rows: after the contacts */
            OracleJoint* j = &g_joints[jg].jt[n];
            Vec3 pA = joint_to_world(w, j->ta, j->ia, j->la), pB = joint_to_world(w, j->tb, j->ib, j->lb);
            BodyRef a = body_ref(w, j->ta, j->ia, pA), b = body_ref(w, j->tb == -1 ? 2 : j->tb, j->ib, pB);
            const int rows = j->kind == ORACLE_JOINT_ROD ? 1 : 3;
            for (int r = 0; r < rows; r++) {
                Vec3 e = { r == 0 ? 1.0f : 0.0f, r == 1 ? 1.0f : 0.0f, r == 2 ? 1.0f : 0.0f };
                if (j->kind == ORACLE_JOINT_ROD) e = j->dir;
                Vec3 velA = a.dyn ? Vec3Add(body_v(w, &a), Vec3Cross(body_om(w, &a), j->rA)) : (Vec3){0, 0, 0};
                Vec3 velB = b.dyn ? Vec3Add(body_v(w, &b), Vec3Cross(body_om(w, &b), j->rB)) : (Vec3){0, 0, 0};
                float vn = Vec3Dot(Vec3Sub(velB, velA), e);
                float jn = j->mass[r] * (-(vn + 0.0f) + j->bias[r]), old = j->imp[r];
                j->imp[r] = j->imp[r] + jn;                                                  /* bilateral: no clamp, no restitution, no friction */
                jn = j->imp[r] - old;
                Vec3 P = Vec3Scale(e, jn), Pm = Vec3Scale(P, -1.0f);
                push(w, &a, Pm, Vec3Cross(j->rA, Pm));
                push(w, &b, P, Vec3Cross(j->rB, P));
            }
        }
    }
}

/* ===== stages 5-7: IntegratePositions :1928-1977, UpdateSleepStates :1696-1784, Cleanup :2305 */
static void sleep_update(Vec3* v, Vec3* om, bool* sleeping, float* timer, PhysicsWorld* w)
{
    bool low = Vec3Length(*v) < w->settings.sleep_linear_threshold && Vec3Length(*om) < w->settings.sleep_angular_threshold;
    if (low) {
        *timer += w->dt;
        if (*timer >= w->settings.sleep_time_required && !*sleeping) { *sleeping = true; *v = (Vec3){0, 0, 0}; *om = (Vec3){0, 0, 0}; }
    } else { *timer = 0.0f; *sleeping = false; }
}
static void stage_positions_sleep(PhysicsWorld* w)
{
    float dt = w->dt; SphereSystem* s = w->spheres; BoxSystem* b = w->boxes;
    for (int i = 0; i < s->count; i++) {
        if (s->is_static[i]) continue;
        if (!s->is_sleeping[i]) s->positions[i] = Vec3Add(s->positions[i], Vec3Scale(s->velocities[i], dt));
    }
    for (int i = 0; i < b->count; i++) {
        if (b->is_static[i] || b->is_sleeping[i]) continue;
        b->positions[i] = Vec3Add(b->positions[i], Vec3Scale(b->velocities[i], dt));
        Vec3 om = b->angular_velocities[i]; Quat q = b->rotations[i];
        Quat qd = QuatMultiply((Quat){om.x, om.y, om.z, 0.0f}, q);
        qd.x *= 0.5f; qd.y *= 0.5f; qd.z *= 0.5f; qd.w *= 0.5f;
        b->rotations[i] = QuatNormalize((Quat){ q.x + qd.x * dt, q.y + qd.y * dt, q.z + qd.z * dt, q.w + qd.w * dt });
    }
    for (int i = 0; i < s->count; i++) if (!s->is_static[i]) sleep_update(&s->velocities[i], &s->angular_velocities[i], &s->is_sleeping[i], &s->sleep_timers[i], w);
    for (int i = 0; i < b->count; i++) if (!b->is_static[i]) sleep_update(&b->velocities[i], &b->angular_velocities[i], &b->is_sleeping[i], &b->sleep_timers[i], w);
}

static void oracle_step(PhysicsWorld* w, float dt, int pruned)
{
    if (!w) return;
    if (dt <= 0.0f || dt < 1e-6f) return;                                                    /* :1791 */
    w->dt = dt;
    stage_springs(w);
    stage_forces_velocities(w);
    stage_detect(w, pruned);
    stage_solve(w);
    stage_positions_sleep(w);
    w->collisions->count = 0;                                                                /* :2310 */
}
void PhysicsStep(PhysicsWorld* w, float dt) { oracle_step(w, dt, 0); }
void OraclePhysicsStepPruned(PhysicsWorld* w, float dt) { oracle_step(w, dt, 1); }
const char* BallboxBackendName(void) { return "oracle/c-restatement"; }
