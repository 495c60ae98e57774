/*
 * ballbox_host.c -- the C host library behind ballbox.h.
 *
 * Owns the public structs (identical layouts) and the host SoA
 * arrays users read and write directly; forwards PhysicsStep to the C-ABI CUDA
 * layer (bbx_cuda.h).  There is NO CPU implementation of the step in this
 * library: without a CUDA device PhysicsStep records an error and prints it.
 *
 * Reference for behaviour: world/body management ballbox.h:301-445, :642-818,
 * :1629-1693; settings :2483-2544; error convention SURVEY.md section 8b
 * (NULL / -1 / silent return).
 *
 * Every world is a batch of n >= 1 worlds internally: the host arrays of all
 * worlds of a batch are slices of one pinned allocation, so an upload is one
 * DMA per array whatever the number of worlds.
 */
#include <stdint.h>
#include "ballbox.h"
#include "bbx_cuda.h"

#define BBX_MAGIC 0xBA11B0C5u

struct BbxWorldPriv;

struct PhysicsWorldBatch {
    uint32_t magic;
    int n_worlds, cap_s, cap_b, cap_c;
    int single;                         /* created by CreatePhysicsWorld */
    struct BbxWorldPriv* worlds;
    /* one allocation per field, n_worlds * cap records */
    BbxHostBodies hb;
    int *s_count, *b_count;
    OnCollision *s_cb, *b_cb;
    ContactConstraint* constraints;     /* n_worlds * cap_c */
    int* ccount;
    int pinned[32];
    int pinned_constraints;
    BbxDev* dev;
    int device;
    void* stream;
    int sync_mode, schedule;
    int sdf_id; float sdf_params[8];
    /* baked host SDF (ballbox_ext.h): region, resolution, and which function the device grid holds */
    int bake_region_set, bake_res, bake_dirty;
    float warm_start;                           /* BallboxSetWarmStarting */
    int bake_no_gradient; float baked_eps;      /* gradient channels are sampled with the sdf_normal_epsilon of bake time */
    float bake_lo[3], bake_hi[3];
    SDFFunction baked_fn;
    int uploaded, forces_dirty, n_callbacks, cb_dirty, timing;
    int dev_ahead;                      /* device steps ran since the host arrays were last in sync (upload or body download) */
    int *wake_ids; int n_wake, cap_wake;    /* bodies woken by Add*Force/Torque while the state is resident (:420-423) */
    CollisionContact* ev_contacts; int *ev_steps, *ev_worlds; int ev_cap;
    /* joints and springs (extension, ballbox_ext.h): creation order, body slots of bbx_cuda.h */
    BbxJointDef* joints; int n_joints, cap_joints;
    BbxSpringDef* springs; int n_springs, cap_springs;
    int joints_dirty;                   /* the device lists are out of date */
    int has_err; char err[256];
};

typedef struct BbxWorldPriv {
    PhysicsWorld pub;                   /* MUST be first: users hold &pub */
    uint32_t magic;
    struct PhysicsWorldBatch* batch;
    int index;
    SphereSystem spheres;
    BoxSystem boxes;
    CollisionCollection collisions;
    ConstraintCollection constraints;
    int n_joints, n_springs;            /* this world's share of the batch's joint / spring lists (extension) */
} BbxWorldPriv;

/* ------------------------------------------------------------------------- */
/* small math, reference ballbox.h:501-639                                   */
/* ------------------------------------------------------------------------- */
Vec3 Vec3Add(Vec3 a, Vec3 b) { Vec3 r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; return r; }
Vec3 Vec3Sub(Vec3 a, Vec3 b) { Vec3 r; r.x = a.x - b.x; r.y = a.y - b.y; r.z = a.z - b.z; return r; }
Vec3 Vec3Scale(Vec3 v, float s) { Vec3 r; r.x = v.x * s; r.y = v.y * s; r.z = v.z * s; return r; }
float Vec3Dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
Vec3 Vec3Cross(Vec3 a, Vec3 b)
{
    Vec3 r;
    r.x = a.y * b.z - a.z * b.y;
    r.y = a.z * b.x - a.x * b.z;
    r.z = a.x * b.y - a.y * b.x;
    return r;
}
float Vec3Length(Vec3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
Vec3 Vec3Normalize(Vec3 v)
{
    float len = Vec3Length(v);
    Vec3 zero = {0.0f, 0.0f, 0.0f};
    return (len < 1e-6f) ? zero : Vec3Scale(v, 1.0f / len);
}
Quat QuatIdentity(void) { Quat q = {0.0f, 0.0f, 0.0f, 1.0f}; return q; }
Quat QuatMultiply(Quat a, Quat b)
{
    Quat r;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return r;
}
Quat QuatNormalize(Quat q)
{
    float len = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    if (len == 0.0f) return QuatIdentity();
    Quat r = { q.x / len, q.y / len, q.z / len, q.w / len };
    return r;
}
Quat QuatConjugate(Quat q) { Quat r = { -q.x, -q.y, -q.z, q.w }; return r; }
Quat QuatFromAxisAngle(Vec3 axis, float angle)
{
    Vec3 n = Vec3Scale(axis, 1.0f / sqrtf(Vec3Dot(axis, axis)));
    float half = angle * 0.5f;
    float s = sinf(half);
    Quat r = { n.x * s, n.y * s, n.z * s, cosf(half) };
    return r;
}
Quat QuatFromEuler(float pitch, float yaw, float roll)
{
    float hp = pitch * 0.5f, hy = yaw * 0.5f, hr = roll * 0.5f;
    float cp = cosf(hp), sp = sinf(hp), cy = cosf(hy), sy = sinf(hy), cr = cosf(hr), sr = sinf(hr);
    Quat r;
    r.x = sp * cy * cr - cp * sy * sr;
    r.y = cp * sy * cr + sp * cy * sr;
    r.z = cp * cy * sr - sp * sy * cr;
    r.w = cp * cy * cr + sp * sy * sr;
    return r;
}
void QuatToAxisAngle(Quat q, Vec3* axis, float* angle)
{
    q = QuatNormalize(q);
    if (fabsf(q.w) >= 1.0f) { *angle = 0.0f; axis->x = 0.0f; axis->y = 1.0f; axis->z = 0.0f; return; }
    *angle = 2.0f * acosf(q.w);
    float sh = sqrtf(1.0f - q.w * q.w);
    if (sh < 1e-6f) { axis->x = 0.0f; axis->y = 1.0f; axis->z = 0.0f; }
    else { axis->x = q.x / sh; axis->y = q.y / sh; axis->z = q.z / sh; }
}
Vec3 QuatRotateVec3(Quat q, Vec3 v)
{
    Vec3 u = { q.x, q.y, q.z };
    float s = q.w;
    Vec3 a = Vec3Scale(v, s * s - Vec3Dot(u, u));
    Vec3 b = Vec3Scale(u, 2.0f * Vec3Dot(u, v));
    Vec3 c = Vec3Scale(Vec3Cross(u, v), 2.0f * s);
    return Vec3Add(Vec3Add(a, b), c);
}

/* ------------------------------------------------------------------------- */
/* helpers                                                                   */
/* ------------------------------------------------------------------------- */
static BbxWorldPriv* priv_of(PhysicsWorld* w)
{
    BbxWorldPriv* p = (BbxWorldPriv*)w;
    return (p && p->magic == BBX_MAGIC) ? p : NULL;
}

static void batch_fail(struct PhysicsWorldBatch* b, const char* msg)
{
    if (!b || b->has_err) return;
    snprintf(b->err, sizeof(b->err), "%s", msg);
    b->has_err = 1;
    fprintf(stderr, "ballbox-b200: %s\n", msg);
}

static int batch_check_dev(struct PhysicsWorldBatch* b, int rc)
{
    if (rc == 0) return 0;
    const char* e = b->dev ? bbx_dev_error(b->dev) : NULL;
    batch_fail(b, e ? e : "CUDA layer call failed");
    return rc;
}

static void* arr_alloc(struct PhysicsWorldBatch* b, int slot, size_t bytes) { return bbx_host_alloc(bytes, &b->pinned[slot]); }

/* ------------------------------------------------------------------------- */
/* batch / world lifetime                                                    */
/* ------------------------------------------------------------------------- */
static void settings_defaults(PhysicsSettings* s)          /* values: reference ballbox.h:2487-2514 */
{
    s->linear_damping = 0.98f;  s->angular_damping = 0.98f;
    s->restitution = 0.0f;      s->friction = 0.1f;       s->solver_iterations = 25;
    s->baumgarte_bias = 0.3f;   s->allowed_penetration = 0.05f;  s->velocity_threshold = 0.02f;
    s->sleep_linear_threshold = 0.1f; s->sleep_angular_threshold = 0.1f; s->sleep_time_required = 0.05f;
    s->manifold_contact_tolerance = 0.01f; s->manifold_penetration_tolerance = -0.5f; s->manifold_max_contacts = 8;
    s->collision_epsilon = 1e-6f; s->sdf_normal_epsilon = 0.001f;
    s->vector_normalize_epsilon = 1e-6f; s->quaternion_epsilon = 1e-6f;
}

static void batch_free(struct PhysicsWorldBatch* b)
{
    free(b->joints); free(b->springs);
    if (!b) return;
    if (b->dev) bbx_dev_destroy(b->dev);
    void* arrs[] = { b->hb.s_pos, b->hb.s_radius, b->hb.s_vel, b->hb.s_force, b->hb.s_mass, b->hb.s_angvel,
                     b->hb.s_torque, b->hb.s_inertia, b->hb.s_timer, b->hb.s_static, b->hb.s_sleep, b->hb.s_hascb,
                     b->hb.b_pos, b->hb.b_size, b->hb.b_rot, b->hb.b_vel, b->hb.b_force, b->hb.b_mass, b->hb.b_angvel,
                     b->hb.b_torque, b->hb.b_inertia, b->hb.b_timer, b->hb.b_static, b->hb.b_sleep, b->hb.b_hascb };
    for (int i = 0; i < 25; i++) bbx_host_free(arrs[i], b->pinned[i]);
    bbx_host_free(b->constraints, b->pinned_constraints);
    if (b->worlds) for (int w = 0; w < b->n_worlds; w++) free(b->worlds[w].collisions.contacts);
    free(b->ev_contacts); free(b->ev_steps); free(b->ev_worlds); free(b->wake_ids);
    free(b->s_count); free(b->b_count); free(b->s_cb); free(b->b_cb); free(b->ccount); free(b->worlds);
    b->magic = 0;
    free(b);
}

static struct PhysicsWorldBatch* batch_create(int n, int max_s, int max_b, int max_c, int single)
{
    if (n < 1 || max_s < 0 || max_b < 0 || max_c < 0) return NULL;
    struct PhysicsWorldBatch* b = (struct PhysicsWorldBatch*)calloc(1, sizeof(*b));
    if (!b) return NULL;
    b->magic = BBX_MAGIC; b->n_worlds = n; b->cap_s = max_s; b->cap_b = max_b; b->cap_c = max_c; b->single = single;
    size_t NS = (size_t)n * max_s, NB = (size_t)n * max_b;
    BbxHostBodies* h = &b->hb;
    h->n_worlds = n; h->cap_s = max_s; h->cap_b = max_b;
    int k = 0;
    h->s_pos = arr_alloc(b, k++, 12 * NS);    h->s_radius = arr_alloc(b, k++, 4 * NS);  h->s_vel = arr_alloc(b, k++, 12 * NS);
    h->s_force = arr_alloc(b, k++, 12 * NS);  h->s_mass = arr_alloc(b, k++, 4 * NS);    h->s_angvel = arr_alloc(b, k++, 12 * NS);
    h->s_torque = arr_alloc(b, k++, 12 * NS); h->s_inertia = arr_alloc(b, k++, 4 * NS); h->s_timer = arr_alloc(b, k++, 4 * NS);
    h->s_static = arr_alloc(b, k++, NS);      h->s_sleep = arr_alloc(b, k++, NS);       h->s_hascb = arr_alloc(b, k++, NS);
    h->b_pos = arr_alloc(b, k++, 12 * NB);    h->b_size = arr_alloc(b, k++, 12 * NB);   h->b_rot = arr_alloc(b, k++, 16 * NB);
    h->b_vel = arr_alloc(b, k++, 12 * NB);    h->b_force = arr_alloc(b, k++, 12 * NB);  h->b_mass = arr_alloc(b, k++, 4 * NB);
    h->b_angvel = arr_alloc(b, k++, 12 * NB); h->b_torque = arr_alloc(b, k++, 12 * NB); h->b_inertia = arr_alloc(b, k++, 12 * NB);
    h->b_timer = arr_alloc(b, k++, 4 * NB);   h->b_static = arr_alloc(b, k++, NB);      h->b_sleep = arr_alloc(b, k++, NB);
    h->b_hascb = arr_alloc(b, k++, NB);
    b->s_count = (int*)calloc((size_t)n, sizeof(int)); b->b_count = (int*)calloc((size_t)n, sizeof(int));
    b->ccount = (int*)calloc((size_t)n, sizeof(int));
    h->s_count = b->s_count; h->b_count = b->b_count;
    b->s_cb = (OnCollision*)calloc(NS ? NS : 1, sizeof(OnCollision));
    b->b_cb = (OnCollision*)calloc(NB ? NB : 1, sizeof(OnCollision));
    b->constraints = (ContactConstraint*)bbx_host_alloc(sizeof(ContactConstraint) * (size_t)n * (size_t)(max_c ? max_c : 1),
                                                        &b->pinned_constraints);
    b->worlds = (BbxWorldPriv*)calloc((size_t)n, sizeof(BbxWorldPriv));
    int ok = h->s_pos && h->s_radius && h->s_vel && h->s_force && h->s_mass && h->s_angvel && h->s_torque && h->s_inertia &&
             h->s_timer && h->s_static && h->s_sleep && h->s_hascb && h->b_pos && h->b_size && h->b_rot && h->b_vel &&
             h->b_force && h->b_mass && h->b_angvel && h->b_torque && h->b_inertia && h->b_timer && h->b_static &&
             h->b_sleep && h->b_hascb && b->s_count && b->b_count && b->ccount && b->s_cb && b->b_cb && b->constraints && b->worlds;
    if (!ok) { batch_free(b); return NULL; }

    /* constructor defaults of the reference (ballbox.h:327-340, :670-684) */
    for (size_t i = 0; i < NS; i++) { h->s_radius[i] = 1.0f; h->s_mass[i] = 1.0f; h->s_inertia[i] = 1.0f; }
    for (size_t i = 0; i < NB; i++) {
        h->b_size[3 * i] = h->b_size[3 * i + 1] = h->b_size[3 * i + 2] = 1.0f;
        h->b_rot[4 * i + 3] = 1.0f; h->b_mass[i] = 1.0f;
        h->b_inertia[3 * i] = h->b_inertia[3 * i + 1] = h->b_inertia[3 * i + 2] = 1.0f / 6.0f;
    }
    for (int w = 0; w < n; w++) {
        BbxWorldPriv* p = &b->worlds[w];
        p->magic = BBX_MAGIC; p->batch = b; p->index = w;
        size_t so = (size_t)w * max_s, bo = (size_t)w * max_b;
        SphereSystem* s = &p->spheres;
        s->positions = (Vec3*)h->s_pos + so; s->radii = h->s_radius + so; s->velocities = (Vec3*)h->s_vel + so;
        s->forces = (Vec3*)h->s_force + so; s->masses = h->s_mass + so; s->angular_velocities = (Vec3*)h->s_angvel + so;
        s->torques = (Vec3*)h->s_torque + so; s->inertias = h->s_inertia + so; s->is_static = (bool*)h->s_static + so;
        s->is_sleeping = (bool*)h->s_sleep + so; s->sleep_timers = h->s_timer + so; s->collision_callbacks = b->s_cb + so;
        s->count = 0; s->capacity = max_s;
        BoxSystem* x = &p->boxes;
        x->positions = (Vec3*)h->b_pos + bo; x->sizes = (Vec3*)h->b_size + bo; x->rotations = (Quat*)h->b_rot + bo;
        x->velocities = (Vec3*)h->b_vel + bo; x->forces = (Vec3*)h->b_force + bo; x->masses = h->b_mass + bo;
        x->angular_velocities = (Vec3*)h->b_angvel + bo; x->torques = (Vec3*)h->b_torque + bo;
        x->inertias = (Vec3*)h->b_inertia + bo; x->is_static = (bool*)h->b_static + bo;
        x->is_sleeping = (bool*)h->b_sleep + bo; x->sleep_timers = h->b_timer + bo; x->collision_callbacks = b->b_cb + bo;
        x->count = 0; x->capacity = max_b;
        p->collisions.contacts = (CollisionContact*)calloc((size_t)(max_c ? max_c : 1), sizeof(CollisionContact));
        p->collisions.count = 0; p->collisions.capacity = max_c;
        p->constraints.constraints = b->constraints + (size_t)w * max_c;
        p->constraints.count = 0; p->constraints.capacity = max_c;
        if (!p->collisions.contacts) { batch_free(b); return NULL; }
        PhysicsWorld* pw = &p->pub;
        pw->spheres = s; pw->boxes = x; pw->collisions = &p->collisions; pw->constraints = &p->constraints;
        pw->gravity.x = pw->gravity.y = pw->gravity.z = 0.0f;         /* ballbox.h:1653 */
        pw->dt = 1.0f / 60.0f; pw->usesSDF = false; pw->SDFCollider = NULL;
        settings_defaults(&pw->settings);
    }
    return b;
}

PhysicsWorld* CreatePhysicsWorld(int max_spheres, int max_boxes, int max_collisions)
{
    struct PhysicsWorldBatch* b = batch_create(1, max_spheres, max_boxes, max_collisions, 1);
    return b ? &b->worlds[0].pub : NULL;
}

void DestroyPhysicsWorld(PhysicsWorld* world)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return;
    if (p->batch->single) batch_free(p->batch);      /* worlds of a real batch die with the batch */
}

PhysicsWorldBatch* CreatePhysicsWorldBatch(int n_worlds, int max_spheres, int max_boxes, int max_collisions)
{
    return batch_create(n_worlds, max_spheres, max_boxes, max_collisions, 0);
}
void DestroyPhysicsWorldBatch(PhysicsWorldBatch* b) { if (b && b->magic == BBX_MAGIC) batch_free(b); }
int  BatchWorldCount(PhysicsWorldBatch* b) { return (b && b->magic == BBX_MAGIC) ? b->n_worlds : 0; }
PhysicsWorld* BatchWorld(PhysicsWorldBatch* b, int i)
{
    if (!b || b->magic != BBX_MAGIC || i < 0 || i >= b->n_worlds) return NULL;
    return &b->worlds[i].pub;
}

void UseSDF(PhysicsWorld* world, SDFFunction fn)                      /* ballbox.h:1688-1693 */
{
    if (!world) return;
    world->SDFCollider = fn;
    world->usesSDF = (fn != NULL);
}

static int bake_sdf(struct PhysicsWorldBatch* b);

void BallboxSetSDFBakeRegion(PhysicsWorld* world, Vec3 lo, Vec3 hi, int resolution)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return;
    struct PhysicsWorldBatch* b = p->batch;
    b->bake_region_set = 1; b->bake_res = resolution; b->bake_dirty = 1;
    b->bake_lo[0] = lo.x; b->bake_lo[1] = lo.y; b->bake_lo[2] = lo.z;
    b->bake_hi[0] = hi.x; b->bake_hi[1] = hi.y; b->bake_hi[2] = hi.z;
}

void BallboxSetSDFBakeGradient(PhysicsWorld* world, int on)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return;
    p->batch->bake_no_gradient = on ? 0 : 1; p->batch->bake_dirty = 1;
}

int BallboxBakeSDF(PhysicsWorld* world)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return BBX_ERR_BAD_ARGUMENT;
    if (p->batch->has_err) return BBX_ERR_STICKY;
    return bake_sdf(p->batch);
}

int BallboxSdfInUse(PhysicsWorld* world)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return BBX_SDF_NONE;
    PhysicsWorld* w0 = &p->batch->worlds[0].pub;
    if (!w0->usesSDF) return BBX_SDF_NONE;
    if (p->batch->sdf_id != BBX_SDF_NONE) return p->batch->sdf_id;
    return w0->SDFCollider ? BBX_SDF_BAKED : BBX_SDF_NONE;
}

void UseSDFDevice(PhysicsWorld* world, int sdf_id, const float* params8)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return;
    p->batch->sdf_id = sdf_id;
    for (int i = 0; i < 8; i++) p->batch->sdf_params[i] = params8 ? params8[i] : 0.0f;
    world->usesSDF = (sdf_id != BBX_SDF_NONE) || (world->SDFCollider != NULL);
}

/* ------------------------------------------------------------------------- */
/* settings, reference ballbox.h:2483-2544                                   */
/* ------------------------------------------------------------------------- */
void ResetPhysicsSettingsToDefaults(PhysicsWorld* w) { if (w) settings_defaults(&w->settings); }
void SetDamping(PhysicsWorld* w, float l, float a) { if (!w) return; w->settings.linear_damping = l; w->settings.angular_damping = a; }
void SetRestitution(PhysicsWorld* w, float r) { if (w) w->settings.restitution = r; }
void SetCorrectionParameters(PhysicsWorld* w, float bias, float allowed)
{
    if (!w) return;
    w->settings.baumgarte_bias = bias; w->settings.allowed_penetration = allowed;
}
void SetSolverIterations(PhysicsWorld* w, int it) { if (w) w->settings.solver_iterations = it; }
void SetManifoldSettings(PhysicsWorld* w, float tol, float pen_tol, int max_contacts)
{
    if (!w) return;
    w->settings.manifold_contact_tolerance = tol;
    w->settings.manifold_penetration_tolerance = pen_tol;
    w->settings.manifold_max_contacts = max_contacts > MAX_MANIFOLD_CONTACTS ? MAX_MANIFOLD_CONTACTS : max_contacts;
}

/* ------------------------------------------------------------------------- */
/* bodies, reference ballbox.h:363-445, :708-818                             */
/* ------------------------------------------------------------------------- */
static const Vec3 V0 = {0.0f, 0.0f, 0.0f};

static int batch_download(struct PhysicsWorldBatch* b, int what);

/* RESIDENT coherence.  While the state lives in HBM the host arrays are stale after the first device step.  An edit
 * of the host arrays that the device must see (a new body, a mass, a static flag, a sync-mode change) therefore
 * first brings the host arrays up to date (one body download, only if device steps ran since the last sync) and then
 * schedules a full re-upload before the next step.  Without this a re-upload would silently snap the simulation back
 * to the last downloaded state. */
static void host_edit_begin(PhysicsWorld* w)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return;
    struct PhysicsWorldBatch* b = p->batch;
    if (b->dev && b->uploaded && b->dev_ahead && !b->has_err) batch_download(b, BBX_DL_BODIES);
}
static void host_edit_end(PhysicsWorld* w)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return;
    p->batch->s_count[p->index] = p->spheres.count;
    p->batch->b_count[p->index] = p->boxes.count;
    p->batch->uploaded = 0;             /* host arrays are the truth again: re-upload before the next step */
}

int AddSphere(PhysicsWorld* world, Vec3 position, float radius)
{
    if (!world || !world->spheres) return -1;
    SphereSystem* s = world->spheres;
    if (s->count >= s->capacity) return -1;
    host_edit_begin(world);
    int i = s->count;
    s->positions[i] = position; s->radii[i] = radius;
    s->velocities[i] = V0; s->forces[i] = V0; s->angular_velocities[i] = V0; s->torques[i] = V0;
    s->masses[i] = 1.0f;
    s->inertias[i] = (2.0f / 5.0f) * s->masses[i] * radius * radius;
    s->count++;
    host_edit_end(world);
    return i;
}

/* ballbox.h:382-386 / :747-752: plain writes into a system's arrays (no world argument).  COMPAT worlds see them at
 * the next PhysicsStep like any direct array write; a RESIDENT world needs BallboxUpload, as for every host-side write. */
void UpdateSphere(SphereSystem* system, int index, Vec3 position)
{
    if (!system || index < 0 || index >= system->count) return;
    system->positions[index] = position;
}
void UpdateBox(BoxSystem* system, int index, Vec3 position, Quat rotation)
{
    if (!system || index < 0 || index >= system->count) return;
    system->positions[index] = position;
    system->rotations[index] = rotation;
}

void SetSphereMass(PhysicsWorld* world, int i, float mass)
{
    if (!world || !world->spheres) return;
    SphereSystem* s = world->spheres;
    if (i < 0 || i >= s->count || mass <= 0.0f) return;
    host_edit_begin(world);
    s->masses[i] = mass;
    float r = s->radii[i];
    s->inertias[i] = (2.0f / 5.0f) * mass * r * r;
    host_edit_end(world);
}

void SetSphereStatic(PhysicsWorld* world, int i, bool is_static)
{
    if (!world || !world->spheres) return;
    SphereSystem* s = world->spheres;
    if (i < 0 || i >= s->count) return;
    host_edit_begin(world);
    s->is_static[i] = is_static;
    if (is_static) { s->velocities[i] = V0; s->angular_velocities[i] = V0; s->forces[i] = V0; s->torques[i] = V0; }
    host_edit_end(world);
}

/* Add*Force/Torque wake a sleeping body even when the force is zero (:420-423, :434-437, :793-796, :807-810).  The
 * host's is_sleeping is stale while the state is resident, so the body is ALSO put on a wake list that the next
 * resident step applies on the device before it integrates (bbx_dev_wake_bodies). */
static void mark_forces(PhysicsWorld* w, int is_box, int i)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return;
    struct PhysicsWorldBatch* b = p->batch;
    b->forces_dirty = 1;
    if (!b->uploaded) return;                          /* the next step uploads the host flags themselves */
    if (b->n_wake == b->cap_wake) {
        int cap = b->cap_wake ? 2 * b->cap_wake : 256;
        int* q = (int*)realloc(b->wake_ids, sizeof(int) * (size_t)cap);
        if (!q) return;
        b->wake_ids = q; b->cap_wake = cap;
    }
    b->wake_ids[b->n_wake++] = is_box ? b->n_worlds * b->cap_s + p->index * b->cap_b + i : p->index * b->cap_s + i;
}

void AddSphereForce(PhysicsWorld* world, int i, Vec3 f)
{
    if (!world || !world->spheres) return;
    SphereSystem* s = world->spheres;
    if (i < 0 || i >= s->count) return;
    if (s->is_sleeping[i]) { s->is_sleeping[i] = false; s->sleep_timers[i] = 0.0f; }
    s->forces[i] = Vec3Add(s->forces[i], f);
    mark_forces(world, 0, i);
}

void AddSphereTorque(PhysicsWorld* world, int i, Vec3 t)
{
    if (!world || !world->spheres) return;
    SphereSystem* s = world->spheres;
    if (i < 0 || i >= s->count) return;
    if (s->is_sleeping[i]) { s->is_sleeping[i] = false; s->sleep_timers[i] = 0.0f; }
    s->torques[i] = Vec3Add(s->torques[i], t);
    mark_forces(world, 0, i);
}

static Vec3 box_inertia(float m, Vec3 half)
{
    float w = half.x * 2.0f, h = half.y * 2.0f, d = half.z * 2.0f;
    Vec3 I = { (m / 12.0f) * (h * h + d * d), (m / 12.0f) * (w * w + d * d), (m / 12.0f) * (w * w + h * h) };
    return I;
}

int AddBox(PhysicsWorld* world, Vec3 position, Vec3 size, Quat rotation)
{
    if (!world || !world->boxes) return -1;
    BoxSystem* b = world->boxes;
    if (b->count >= b->capacity) return -1;
    host_edit_begin(world);
    int i = b->count;
    b->positions[i] = position; b->sizes[i] = size; b->rotations[i] = rotation;
    b->velocities[i] = V0; b->forces[i] = V0; b->angular_velocities[i] = V0; b->torques[i] = V0;
    b->masses[i] = 1.0f;
    b->inertias[i] = box_inertia(b->masses[i], size);
    b->is_static[i] = false; b->is_sleeping[i] = false; b->sleep_timers[i] = 0.0f;
    b->count++;
    host_edit_end(world);
    return i;
}

void SetBoxMass(PhysicsWorld* world, int i, float mass)
{
    if (!world || !world->boxes) return;
    BoxSystem* b = world->boxes;
    if (i < 0 || i >= b->count || mass <= 0.0f) return;
    host_edit_begin(world);
    b->masses[i] = mass;
    b->inertias[i] = box_inertia(mass, b->sizes[i]);
    host_edit_end(world);
}

void SetBoxStatic(PhysicsWorld* world, int i, bool is_static)
{
    if (!world || !world->boxes) return;
    BoxSystem* b = world->boxes;
    if (i < 0 || i >= b->count) return;
    host_edit_begin(world);
    b->is_static[i] = is_static;
    if (is_static) { b->velocities[i] = V0; b->angular_velocities[i] = V0; b->forces[i] = V0; b->torques[i] = V0; }
    host_edit_end(world);
}

void AddBoxForce(PhysicsWorld* world, int i, Vec3 f)
{
    if (!world || !world->boxes) return;
    BoxSystem* b = world->boxes;
    if (i < 0 || i >= b->count) return;
    if (b->is_sleeping[i]) { b->is_sleeping[i] = false; b->sleep_timers[i] = 0.0f; }
    b->forces[i] = Vec3Add(b->forces[i], f);
    mark_forces(world, 1, i);
}

void AddBoxTorque(PhysicsWorld* world, int i, Vec3 t)
{
    if (!world || !world->boxes) return;
    BoxSystem* b = world->boxes;
    if (i < 0 || i >= b->count) return;
    if (b->is_sleeping[i]) { b->is_sleeping[i] = false; b->sleep_timers[i] = 0.0f; }
    b->torques[i] = Vec3Add(b->torques[i], t);
    mark_forces(world, 1, i);
}

static void recount_callbacks(struct PhysicsWorldBatch* b)
{
    int n = 0;
    size_t NS = (size_t)b->n_worlds * b->cap_s, NB = (size_t)b->n_worlds * b->cap_b;
    for (size_t i = 0; i < NS; i++) { b->hb.s_hascb[i] = b->s_cb[i] != NULL; n += b->hb.s_hascb[i]; }
    for (size_t i = 0; i < NB; i++) { b->hb.b_hascb[i] = b->b_cb[i] != NULL; n += b->hb.b_hascb[i]; }
    b->n_callbacks = n;
    b->cb_dirty = 1;
}

void SetSphereCollisionCallback(PhysicsWorld* world, int i, OnCollision cb)
{
    if (!world || !world->spheres || i < 0 || i >= world->spheres->count) return;
    world->spheres->collision_callbacks[i] = cb;
    BbxWorldPriv* p = priv_of(world);
    if (p) recount_callbacks(p->batch);
}

void SetBoxCollisionCallback(PhysicsWorld* world, int i, OnCollision cb)
{
    if (!world || !world->boxes || i < 0 || i >= world->boxes->count) return;
    world->boxes->collision_callbacks[i] = cb;
    BbxWorldPriv* p = priv_of(world);
    if (p) recount_callbacks(p->batch);
}

/* standalone systems and constraint collections (API completeness; ballbox.h:301-361, :642-706, :1138-1170) */
SphereSystem* CreateSphereSystem(int capacity)
{
    if (capacity < 0) return NULL;
    SphereSystem* s = (SphereSystem*)calloc(1, sizeof(SphereSystem));
    if (!s) return NULL;
    size_t n = capacity ? (size_t)capacity : 1;
    s->positions = calloc(n, sizeof(Vec3)); s->radii = calloc(n, 4); s->velocities = calloc(n, sizeof(Vec3));
    s->forces = calloc(n, sizeof(Vec3)); s->masses = calloc(n, 4); s->angular_velocities = calloc(n, sizeof(Vec3));
    s->torques = calloc(n, sizeof(Vec3)); s->inertias = calloc(n, 4); s->is_static = calloc(n, 1);
    s->is_sleeping = calloc(n, 1); s->sleep_timers = calloc(n, 4); s->collision_callbacks = calloc(n, sizeof(OnCollision));
    s->capacity = capacity;
    if (!s->positions || !s->radii || !s->velocities || !s->forces || !s->masses || !s->angular_velocities || !s->torques ||
        !s->inertias || !s->is_static || !s->is_sleeping || !s->sleep_timers || !s->collision_callbacks) {
        DestroySphereSystem(s); return NULL;
    }
    for (int i = 0; i < capacity; i++) { s->radii[i] = 1.0f; s->masses[i] = 1.0f; s->inertias[i] = 1.0f; }
    return s;
}
void DestroySphereSystem(SphereSystem* s)
{
    if (!s) return;
    free(s->positions); free(s->radii); free(s->velocities); free(s->forces); free(s->masses); free(s->angular_velocities);
    free(s->torques); free(s->inertias); free(s->is_static); free(s->is_sleeping); free(s->sleep_timers);
    free(s->collision_callbacks); free(s);
}
BoxSystem* CreateBoxSystem(int capacity)
{
    if (capacity < 0) return NULL;
    BoxSystem* b = (BoxSystem*)calloc(1, sizeof(BoxSystem));
    if (!b) return NULL;
    size_t n = capacity ? (size_t)capacity : 1;
    b->positions = calloc(n, sizeof(Vec3)); b->sizes = calloc(n, sizeof(Vec3)); b->rotations = calloc(n, sizeof(Quat));
    b->velocities = calloc(n, sizeof(Vec3)); b->forces = calloc(n, sizeof(Vec3)); b->masses = calloc(n, 4);
    b->angular_velocities = calloc(n, sizeof(Vec3)); b->torques = calloc(n, sizeof(Vec3)); b->inertias = calloc(n, sizeof(Vec3));
    b->is_static = calloc(n, 1); b->is_sleeping = calloc(n, 1); b->sleep_timers = calloc(n, 4);
    b->collision_callbacks = calloc(n, sizeof(OnCollision));
    b->capacity = capacity;
    if (!b->positions || !b->sizes || !b->rotations || !b->velocities || !b->forces || !b->masses || !b->angular_velocities ||
        !b->torques || !b->inertias || !b->is_static || !b->is_sleeping || !b->sleep_timers || !b->collision_callbacks) {
        DestroyBoxSystem(b); return NULL;
    }
    for (int i = 0; i < capacity; i++) {
        Vec3 one = {1.0f, 1.0f, 1.0f}, in = {1.0f / 6.0f, 1.0f / 6.0f, 1.0f / 6.0f};
        b->sizes[i] = one; b->rotations[i] = QuatIdentity(); b->masses[i] = 1.0f; b->inertias[i] = in;
    }
    return b;
}
void DestroyBoxSystem(BoxSystem* b)
{
    if (!b) return;
    free(b->positions); free(b->sizes); free(b->rotations); free(b->velocities); free(b->forces); free(b->masses);
    free(b->angular_velocities); free(b->torques); free(b->inertias); free(b->is_static); free(b->is_sleeping);
    free(b->sleep_timers); free(b->collision_callbacks); free(b);
}
ConstraintCollection* CreateConstraintCollection(int capacity)
{
    if (capacity < 0) return NULL;
    ConstraintCollection* c = (ConstraintCollection*)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->constraints = (ContactConstraint*)calloc(capacity ? (size_t)capacity : 1, sizeof(ContactConstraint));
    if (!c->constraints) { free(c); return NULL; }
    c->capacity = capacity;
    return c;
}
void DestroyConstraintCollection(ConstraintCollection* c) { if (c) { free(c->constraints); free(c); } }
void ClearConstraints(ConstraintCollection* c) { if (c) c->count = 0; }

/* ------------------------------------------------------------------------- */
/* device plumbing                                                           */
/* ------------------------------------------------------------------------- */
static int batch_ensure_dev(struct PhysicsWorldBatch* b)
{
    if (b->has_err) return BBX_ERR_STICKY;
    if (b->dev) return 0;
    int rc = bbx_dev_create(&b->dev, b->device, b->n_worlds, b->cap_s, b->cap_b, b->cap_c);
    if (rc) {
        const char* e = b->dev ? bbx_dev_error(b->dev) : NULL;
        batch_fail(b, e ? e : (rc == BBX_ERR_NO_DEVICE
                               ? "no CUDA device: PhysicsStep needs a B200-class GPU (this library has no CPU path)"
                               : "bbx_dev_create failed"));
        if (b->dev) { bbx_dev_destroy(b->dev); b->dev = NULL; }
        return rc;
    }
    bbx_dev_set_stream(b->dev, b->stream);
    if (b->timing) bbx_dev_set_timing(b->dev, 1);
    return 0;
}

static int batch_upload(struct PhysicsWorldBatch* b)
{
    int rc = batch_ensure_dev(b);
    if (rc) return rc;
    for (int w = 0; w < b->n_worlds; w++) { b->s_count[w] = b->worlds[w].spheres.count; b->b_count[w] = b->worlds[w].boxes.count; }
    rc = batch_check_dev(b, bbx_dev_upload(b->dev, &b->hb));
    if (!rc) { b->uploaded = 1; b->cb_dirty = 0; b->dev_ahead = 0; b->n_wake = 0; }
    return rc;
}

/* Bake world 0's host SDF (reference UseSDF, ballbox.h:1688) into a regular grid on the device.  The GPU
 * cannot call a host function pointer; this is the one-off sampling the step kernels interpolate. */
static int bake_sdf(struct PhysicsWorldBatch* b)
{
    PhysicsWorld* w0 = &b->worlds[0].pub;
    SDFFunction fn = w0->SDFCollider;
    if (!fn) { batch_fail(b, "BallboxBakeSDF: no host SDF was set with UseSDF"); return BBX_ERR_BAD_ARGUMENT; }
    int rc = batch_ensure_dev(b);
    if (rc) return rc;
    float lo[3], hi[3];
    if (b->bake_region_set) {
        for (int a = 0; a < 3; a++) { lo[a] = b->bake_lo[a]; hi[a] = b->bake_hi[a]; }
    } else {
        /* bounding box of every body of every world, grown by 25 % + 2 units on each side */
        int any = 0;
        for (int a = 0; a < 3; a++) { lo[a] = 0.0f; hi[a] = 0.0f; }
        for (int w = 0; w < b->n_worlds; w++) {
            SphereSystem* S = &b->worlds[w].spheres; BoxSystem* B = &b->worlds[w].boxes;
            for (int i = 0; i < S->count + B->count; i++) {
                Vec3 c; float r;
                if (i < S->count) { c = S->positions[i]; r = S->radii[i]; }
                else {
                    int k = i - S->count; Quat q = B->rotations[k];
                    float q2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
                    c = B->positions[k]; r = Vec3Length(B->sizes[k]) * (q2 > 1.0f ? q2 : 1.0f);
                }
                float pc[3] = { c.x, c.y, c.z };
                for (int a = 0; a < 3; a++) {
                    if (!(pc[a] == pc[a]) || !(r == r)) continue;                      /* NaN */
                    if (!any || pc[a] - r < lo[a]) lo[a] = pc[a] - r;
                    if (!any || pc[a] + r > hi[a]) hi[a] = pc[a] + r;
                }
                any = 1;
            }
        }
        if (!any) for (int a = 0; a < 3; a++) { lo[a] = -10.0f; hi[a] = 10.0f; }
        for (int a = 0; a < 3; a++) { float g = 0.25f * (hi[a] - lo[a]) + 2.0f; lo[a] -= g; hi[a] += g; }
    }
    int res = b->bake_res > 0 ? b->bake_res : 192;
    if (res < 2) res = 2;
    if (res > 1024) res = 1024;
    float longest = 0.0f;
    for (int a = 0; a < 3; a++) { if (!(hi[a] > lo[a])) hi[a] = lo[a] + 1.0f; if (hi[a] - lo[a] > longest) longest = hi[a] - lo[a]; }
    int n[3]; float cell;
    for (;;) {
        cell = longest / (float)(res - 1);
        long long total = 1;
        for (int a = 0; a < 3; a++) { n[a] = (int)ceilf((hi[a] - lo[a]) / cell) + 1; if (n[a] < 2) n[a] = 2; total *= n[a]; }
        if (total <= (1ll << 26) || res <= 2) break;
        res = res * 3 / 4;                              /* keep the grid below 64 M nodes (256 MB, 1 GB with gradients) */
    }
    size_t count = (size_t)n[0] * n[1] * n[2];
    const int ch = b->bake_no_gradient ? 1 : 4;
    float* values = (float*)malloc(count * sizeof(float) * (size_t)ch);
    if (!values) { batch_fail(b, "BallboxBakeSDF: out of host memory"); return BBX_ERR_BAD_ARGUMENT; }
    const float eps = w0->settings.sdf_normal_epsilon;
    size_t at = 0;
    for (int k = 0; k < n[2]; k++)
        for (int j = 0; j < n[1]; j++)
            for (int i = 0; i < n[0]; i++) {
                Vec3 p = { lo[0] + (float)i * cell, lo[1] + (float)j * cell, lo[2] + (float)k * cell };
                values[at++] = fn(p);
                if (ch == 4) {
                    /* the three components of SDF_EstimateNormal before normalisation (ballbox.h:2325-2327) */
                    Vec3 ex = { eps, 0.0f, 0.0f }, ey = { 0.0f, eps, 0.0f }, ez = { 0.0f, 0.0f, eps };
                    values[at++] = (fn(Vec3Add(p, ex)) - fn(Vec3Sub(p, ex))) / (2.0f * eps);
                    values[at++] = (fn(Vec3Add(p, ey)) - fn(Vec3Sub(p, ey))) / (2.0f * eps);
                    values[at++] = (fn(Vec3Add(p, ez)) - fn(Vec3Sub(p, ez))) / (2.0f * eps);
                }
            }
    rc = batch_check_dev(b, bbx_dev_set_baked_sdf(b->dev, values, ch, n[0], n[1], n[2], lo, cell));
    free(values);
    if (!rc) { b->baked_fn = fn; b->bake_dirty = 0; b->baked_eps = eps; }
    return rc;
}

static int fill_params(struct PhysicsWorldBatch* b, float dt, int has_forces, int want_callbacks, BbxStepParams* p)
{
    PhysicsWorld* w0 = &b->worlds[0].pub;
    memset(p, 0, sizeof(*p));
    p->gravity[0] = w0->gravity.x; p->gravity[1] = w0->gravity.y; p->gravity[2] = w0->gravity.z;
    p->dt = dt; p->settings = w0->settings;
    p->sdf_id = BBX_SDF_NONE;
    if (w0->usesSDF) {
        if (b->sdf_id == BBX_SDF_NONE) {
            /* reference condition is usesSDF && SDFCollider (ballbox.h:2075); a host pointer cannot run on the GPU:
             * it is baked once into a grid (ballbox_ext.h) */
            if (w0->SDFCollider) {
                if (b->baked_fn != w0->SDFCollider || b->bake_dirty ||
                    (!b->bake_no_gradient && b->baked_eps != w0->settings.sdf_normal_epsilon)) { int rc = bake_sdf(b); if (rc) return rc; }
                p->sdf_id = BBX_SDF_BAKED;
            }
        } else {
            p->sdf_id = b->sdf_id;
            memcpy(p->sdf_params, b->sdf_params, sizeof(p->sdf_params));
        }
    }
    p->solver_schedule = b->schedule;
    p->warm_start = b->warm_start;
    p->has_forces = has_forces;
    p->want_callbacks = want_callbacks;
    return 0;
}

static void zero_host_forces(struct PhysicsWorldBatch* b)
{
    for (int w = 0; w < b->n_worlds; w++) {
        BbxWorldPriv* p = &b->worlds[w];
        memset(p->spheres.forces, 0, sizeof(Vec3) * (size_t)p->spheres.count);
        memset(p->spheres.torques, 0, sizeof(Vec3) * (size_t)p->spheres.count);
        memset(p->boxes.forces, 0, sizeof(Vec3) * (size_t)p->boxes.count);
        memset(p->boxes.torques, 0, sizeof(Vec3) * (size_t)p->boxes.count);
    }
    b->forces_dirty = 0;
}

static int batch_download(struct PhysicsWorldBatch* b, int what)
{
    if (!b->dev) return 0;
    int rc = 0;
    if (what & BBX_DL_BODIES) { rc = batch_check_dev(b, bbx_dev_download(b->dev, &b->hb)); if (!rc) b->dev_ahead = 0; }
    if (!rc && (what & BBX_DL_CONSTRAINTS)) {
        rc = batch_check_dev(b, bbx_dev_download_constraints(b->dev, b->constraints, b->cap_c, b->ccount));
        if (!rc) for (int w = 0; w < b->n_worlds; w++) b->worlds[w].constraints.count = b->ccount[w];
    }
    return rc;
}

/* Replay the collision callbacks in the reference's firing order (SURVEY.md 3.3;
 * ballbox.h:1999-2005, :2034-2044, :2063-2069, :2100-2103).  The constraint list
 * is already in emission order; the contact handed out is a host copy. */
static void replay_callbacks(BbxWorldPriv* p)
{
    ConstraintCollection* cc = &p->constraints;
    SphereSystem* S = &p->spheres; BoxSystem* B = &p->boxes;
    int prev_a = -1, prev_b = -1;
    for (int i = 0; i < cc->count; i++) {
        ContactConstraint* c = &cc->constraints[i];
        CollisionContact k;
        k.bodyA_type = c->bodyA_type; k.bodyA_index = c->bodyA_index; k.bodyB_type = c->bodyB_type; k.bodyB_index = c->bodyB_index;
        k.contact_point = c->contact_point; k.contact_normal = c->contact_normal; k.penetration = c->penetration;
        int a = c->bodyA_index, b = c->bodyB_index;
        if (c->bodyA_type == 0 && c->bodyB_type == 0) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 0, b, &k);
            if (S->collision_callbacks[b]) S->collision_callbacks[b](b, 0, a, &k);
        } else if (c->bodyA_type == 1 && c->bodyB_type == 1) {
            if (a != prev_a || b != prev_b) {                       /* once per pair, first manifold point */
                if (B->collision_callbacks[a]) B->collision_callbacks[a](a, 1, b, &k);
                if (B->collision_callbacks[b]) B->collision_callbacks[b](b, 1, a, &k);
            }
            prev_a = a; prev_b = b;
            continue;
        } else if (c->bodyA_type == 0 && c->bodyB_type == 1) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 1, b, &k);
            if (B->collision_callbacks[b]) B->collision_callbacks[b](b, 0, a, &k);
        } else if (c->bodyA_type == 0 && c->bodyB_type == 2) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 2, 0, &k);
        }                                                           /* box-SDF: no callback in the reference */
        prev_a = -1; prev_b = -1;
    }
}

/* device event buffer -> callbacks, same dispatch rules as replay_callbacks */
static int replay_events(struct PhysicsWorldBatch* b)
{
    if (!b->ev_contacts) {
        b->ev_cap = bbx_dev_event_capacity(b->dev);
        b->ev_contacts = (CollisionContact*)malloc(sizeof(CollisionContact) * (size_t)(b->ev_cap ? b->ev_cap : 1));
        b->ev_steps = (int*)malloc(sizeof(int) * (size_t)(b->ev_cap ? b->ev_cap : 1));
        b->ev_worlds = (int*)malloc(sizeof(int) * (size_t)(b->ev_cap ? b->ev_cap : 1));
        if (!b->ev_contacts || !b->ev_steps || !b->ev_worlds) { batch_fail(b, "out of host memory for callback events"); return BBX_ERR_BAD_ARGUMENT; }
    }
    int n = 0;
    int rc = batch_check_dev(b, bbx_dev_fetch_events(b->dev, b->ev_contacts, b->ev_steps, b->ev_worlds, b->ev_cap, &n));
    for (int e = 0; e < n; e++) {
        CollisionContact* k = &b->ev_contacts[e];
        int w = b->ev_worlds[e];
        if (w < 0 || w >= b->n_worlds) continue;
        SphereSystem* S = &b->worlds[w].spheres; BoxSystem* B = &b->worlds[w].boxes;
        int a = k->bodyA_index, c = k->bodyB_index;
        if (k->bodyA_type == 0 && k->bodyB_type == 0) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 0, c, k);
            if (S->collision_callbacks[c]) S->collision_callbacks[c](c, 0, a, k);
        } else if (k->bodyA_type == 1 && k->bodyB_type == 1) {
            if (B->collision_callbacks[a]) B->collision_callbacks[a](a, 1, c, k);
            if (B->collision_callbacks[c]) B->collision_callbacks[c](c, 1, a, k);
        } else if (k->bodyA_type == 0 && k->bodyB_type == 1) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 1, c, k);
            if (B->collision_callbacks[c]) B->collision_callbacks[c](c, 0, a, k);
        } else if (k->bodyA_type == 0 && k->bodyB_type == 2) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 2, 0, k);
        }
    }
    return rc;
}

static int sync_joints(struct PhysicsWorldBatch* b);
static int batch_step(struct PhysicsWorldBatch* b, float dt, int steps)
{
    if (b->has_err) return BBX_ERR_STICKY;
    if (dt <= 0.0f || dt < 1e-6f) return 0;                        /* ballbox.h:1791 */
    if (steps <= 0) return 0;
    int rc;
    for (int w = 0; w < b->n_worlds; w++) b->worlds[w].pub.dt = dt;
    if (b->joints_dirty) {
        if ((rc = batch_ensure_dev(b))) return rc;
        if ((rc = sync_joints(b))) return rc;
    }
    if (b->sync_mode == BBX_SYNC_COMPAT || b->sync_mode == BBX_SYNC_COMPAT_BODIES) {
        /* drop-in semantics: whatever the user wrote into the host arrays is what the step sees.
         * COMPAT_BODIES leaves the constraint list on the device (it is 6x the size of the body state at
         * 6 contacts per body): world->constraints->count is 0 until BallboxDownload(BBX_DL_CONSTRAINTS). */
        const int full = b->sync_mode == BBX_SYNC_COMPAT;
        for (int s = 0; s < steps; s++) {
            BbxStepParams prm;
            if ((rc = batch_upload(b))) return rc;
            /* the full mode replays callbacks from the downloaded list, the other one from device-side events */
            if ((rc = fill_params(b, dt, 1, (!full && b->n_callbacks > 0) ? 1 : 0, &prm))) return rc;
            if ((rc = batch_check_dev(b, bbx_dev_step(b->dev, &prm, 1)))) return rc;
            /* the step is queued, not finished: clear the force arrays (:1836) as soon as the upload has read them */
            if ((rc = batch_check_dev(b, bbx_dev_wait_upload(b->dev)))) return rc;
            zero_host_forces(b);
            if (full) {
                /* body arrays cross the link while the constraint records are exported and sorted */
                if ((rc = batch_check_dev(b, bbx_dev_download_begin(b->dev, &b->hb)))) return rc;
                rc = batch_download(b, BBX_DL_CONSTRAINTS);
                int rc2 = batch_check_dev(b, bbx_dev_download_end(b->dev));
                if (rc || rc2) return rc ? rc : rc2;
            } else if ((rc = batch_download(b, BBX_DL_BODIES))) return rc;
            for (int w = 0; w < b->n_worlds; w++) b->worlds[w].collisions.count = 0;       /* ballbox.h:2310 */
            if (!full) for (int w = 0; w < b->n_worlds; w++) b->worlds[w].constraints.count = 0;
            if (b->n_callbacks > 0) {
                if (full) for (int w = 0; w < b->n_worlds; w++) replay_callbacks(&b->worlds[w]);
                else if ((rc = replay_events(b))) return rc;
            }
        }
        return 0;
    }
    /* resident */
    if (!b->uploaded && (rc = batch_upload(b))) return rc;
    if (b->cb_dirty) {                       /* Set*CollisionCallback since the last upload */
        if ((rc = batch_check_dev(b, bbx_dev_upload_callback_flags(b->dev, &b->hb)))) return rc;
        b->cb_dirty = 0;
    }
    if (b->n_wake > 0) {                     /* Add*Force/Torque on resident bodies: wake them on the device (:420-423) */
        if ((rc = batch_check_dev(b, bbx_dev_wake_bodies(b->dev, b->wake_ids, b->n_wake)))) return rc;
        b->n_wake = 0;
    }
    b->dev_ahead = 1;
    if (b->n_callbacks > 0) {
        /* Events are appended on the device right after detection and replayed here, after the
         * steps, in (step, emission) order.  Chunks bound the size of the event buffer. */
        int done = 0;
        while (done < steps) {
            int chunk = steps - done < 8 ? steps - done : 8;
            BbxStepParams prm;
            int hf = b->forces_dirty;
            if (hf && (rc = batch_check_dev(b, bbx_dev_upload_forces(b->dev, &b->hb)))) return rc;
            if ((rc = fill_params(b, dt, hf, 1, &prm))) return rc;
            if ((rc = batch_check_dev(b, bbx_dev_step(b->dev, &prm, chunk)))) return rc;
            if (hf) {                            /* the force arrays are pinned: clear them only after the queued copies have read them */
                if ((rc = batch_check_dev(b, bbx_dev_wait_upload(b->dev)))) return rc;
                zero_host_forces(b);
            }
            if ((rc = replay_events(b))) return rc;
            done += chunk;
        }
        return 0;
    }
    BbxStepParams prm;
    int hf = b->forces_dirty;
    if (hf && (rc = batch_check_dev(b, bbx_dev_upload_forces(b->dev, &b->hb)))) return rc;
    if ((rc = fill_params(b, dt, hf, 0, &prm))) return rc;
    rc = batch_check_dev(b, bbx_dev_step(b->dev, &prm, steps));
    if (hf) {                                    /* the force arrays are pinned: clear them only after the queued copies have read them */
        int rc2 = batch_check_dev(b, bbx_dev_wait_upload(b->dev));
        if (!rc) rc = rc2;
        zero_host_forces(b);
    }
    return rc;
}

/* ------------------------------------------------------------------------- */
/* joints and springs (extension; sequential definition: oracle/bbx_oracle.c) */
/* ------------------------------------------------------------------------- */
static int joint_body_slot(BbxWorldPriv* p, int type, int index, int world_allowed, int* slot)
{
    struct PhysicsWorldBatch* b = p->batch;
    if (type == -1) { *slot = -1; return world_allowed; }
    if (type == 0 && index >= 0 && index < p->spheres.count) { *slot = p->index * b->cap_s + index; return 1; }
    if (type == 1 && index >= 0 && index < p->boxes.count) { *slot = b->n_worlds * b->cap_s + p->index * b->cap_b + index; return 1; }
    return 0;
}
/* world point -> body frame: boxes rotate, a sphere is attached at its centre (SphereSystem has no rotations), the world is fixed */
static Vec3 joint_to_local(BbxWorldPriv* p, int type, int index, Vec3 pt)
{
    if (type == 0) return V0;
    if (type == 1) return QuatRotateVec3(QuatConjugate(p->boxes.rotations[index]), Vec3Sub(pt, p->boxes.positions[index]));
    return pt;
}
static Vec3 joint_to_world(BbxWorldPriv* p, int type, int index, Vec3 l)
{
    if (type == 0) return Vec3Add(p->spheres.positions[index], l);
    if (type == 1) return Vec3Add(p->boxes.positions[index], QuatRotateVec3(p->boxes.rotations[index], l));
    return l;
}
static int slot_world(struct PhysicsWorldBatch* b, int slot)
{
    const int NS = b->n_worlds * b->cap_s;
    return slot < NS ? slot / (b->cap_s > 0 ? b->cap_s : 1) : (slot - NS) / (b->cap_b > 0 ? b->cap_b : 1);
}
static int joint_add(PhysicsWorld* world, int kind, int ta, int ia, int tb, int ib, Vec3 anchorA, Vec3 anchorB)
{
    BbxWorldPriv* p = priv_of(world);
    int ga, gb;
    if (!p || !joint_body_slot(p, ta, ia, 0, &ga) || !joint_body_slot(p, tb, ib, 1, &gb)) return -1;
    struct PhysicsWorldBatch* b = p->batch;
    host_edit_begin(world);             /* anchors are taken from the CURRENT positions: a resident world is downloaded first */
    if (b->n_joints == b->cap_joints) {
        int cap = b->cap_joints ? 2 * b->cap_joints : 16;
        BbxJointDef* q = (BbxJointDef*)realloc(b->joints, sizeof(BbxJointDef) * (size_t)cap);
        if (!q) return -1;
        b->joints = q; b->cap_joints = cap;
    }
    BbxJointDef* j = &b->joints[b->n_joints];
    Vec3 la = joint_to_local(p, ta, ia, anchorA), lb = joint_to_local(p, tb, ib, anchorB);
    j->kind = kind; j->ga = ga; j->gb = gb;
    j->la[0] = la.x; j->la[1] = la.y; j->la[2] = la.z; j->lb[0] = lb.x; j->lb[1] = lb.y; j->lb[2] = lb.z;
    j->rest = Vec3Length(Vec3Sub(joint_to_world(p, tb, ib, lb), joint_to_world(p, ta, ia, la)));
    b->joints_dirty = 1;
    b->n_joints++;
    return p->n_joints++;                               /* index among this world's joints, like a single world's creation order */
}
int BallboxAddBallJoint(PhysicsWorld* world, int ta, int ia, int tb, int ib, Vec3 anchor) { return joint_add(world, BBX_JOINT_BALL, ta, ia, tb, ib, anchor, anchor); }
int BallboxAddDistanceJoint(PhysicsWorld* world, int ta, int ia, int tb, int ib, Vec3 anchorA, Vec3 anchorB) { return joint_add(world, BBX_JOINT_ROD, ta, ia, tb, ib, anchorA, anchorB); }
int BallboxAddHingeJoint(PhysicsWorld* world, int ta, int ia, int tb, int ib, Vec3 anchor, Vec3 axis, float half_span)
{
    Vec3 d = Vec3Scale(Vec3Normalize(axis), half_span);
    int first = BallboxAddBallJoint(world, ta, ia, tb, ib, Vec3Sub(anchor, d));
    if (first < 0) return -1;
    return BallboxAddBallJoint(world, ta, ia, tb, ib, Vec3Add(anchor, d)) < 0 ? -1 : first;
}
int BallboxAddSpring(PhysicsWorld* world, int ta, int ia, int tb, int ib, Vec3 anchorA, Vec3 anchorB, float rest, float k, float c)
{
    BbxWorldPriv* p = priv_of(world);
    int ga, gb;
    if (!p || !joint_body_slot(p, ta, ia, 0, &ga) || !joint_body_slot(p, tb, ib, 1, &gb)) return -1;
    struct PhysicsWorldBatch* b = p->batch;
    host_edit_begin(world);
    if (b->n_springs == b->cap_springs) {
        int cap = b->cap_springs ? 2 * b->cap_springs : 16;
        BbxSpringDef* q = (BbxSpringDef*)realloc(b->springs, sizeof(BbxSpringDef) * (size_t)cap);
        if (!q) return -1;
        b->springs = q; b->cap_springs = cap;
    }
    BbxSpringDef* sp = &b->springs[b->n_springs];
    Vec3 la = joint_to_local(p, ta, ia, anchorA), lb = joint_to_local(p, tb, ib, anchorB);
    sp->ga = ga; sp->gb = gb; sp->rest = rest; sp->k = k; sp->c = c;
    sp->la[0] = la.x; sp->la[1] = la.y; sp->la[2] = la.z; sp->lb[0] = lb.x; sp->lb[1] = lb.y; sp->lb[2] = lb.z;
    b->joints_dirty = 1;
    b->n_springs++;
    return p->n_springs++;
}
int BallboxJointCount(PhysicsWorld* world) { BbxWorldPriv* p = priv_of(world); return p ? p->n_joints : 0; }
int BallboxSpringCount(PhysicsWorld* world) { BbxWorldPriv* p = priv_of(world); return p ? p->n_springs : 0; }
void BallboxClearJoints(PhysicsWorld* world)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return;
    struct PhysicsWorldBatch* b = p->batch;
    int n = 0;
    for (int k = 0; k < b->n_joints; k++) if (slot_world(b, b->joints[k].ga) != p->index) b->joints[n++] = b->joints[k];
    b->n_joints = n;
    n = 0;
    for (int k = 0; k < b->n_springs; k++) if (slot_world(b, b->springs[k].ga) != p->index) b->springs[n++] = b->springs[k];
    b->n_springs = n;
    p->n_joints = p->n_springs = 0;
    b->joints_dirty = 1;
}
static int sync_joints(struct PhysicsWorldBatch* b)
{
    int rc = batch_check_dev(b, bbx_dev_set_joints(b->dev, b->joints, b->n_joints));
    if (!rc) rc = batch_check_dev(b, bbx_dev_set_springs(b->dev, b->springs, b->n_springs));
    if (!rc) b->joints_dirty = 0;
    return rc;
}

/* ------------------------------------------------------------------------- */
/* public stage functions, reference ballbox.h:217-229                       */
/* ------------------------------------------------------------------------- */
static struct PhysicsWorldBatch* stage_begin(PhysicsWorld* world, BbxStepParams* prm)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return NULL;
    struct PhysicsWorldBatch* b = p->batch;
    if (b->has_err) return NULL;
    if (b->n_worlds != 1) { batch_fail(b, "stage functions work on single worlds, not on members of a batch"); return NULL; }
    if (batch_upload(b)) return NULL;                       /* host arrays are the truth between stages */
    if (fill_params(b, world->dt, 1, 0, prm)) return NULL;
    return b;
}

void ApplyForces(PhysicsWorld* world)                                /* :1820 */
{
    BbxStepParams prm; struct PhysicsWorldBatch* b = stage_begin(world, &prm);
    if (!b) return;
    if (batch_check_dev(b, bbx_dev_run_stage(b->dev, &prm, BBX_STAGE_APPLY_FORCES))) return;
    batch_check_dev(b, bbx_dev_download_forces(b->dev, &b->hb));
}

void IntegrateVelocities(PhysicsWorld* world)                        /* :1841 */
{
    BbxStepParams prm; struct PhysicsWorldBatch* b = stage_begin(world, &prm);
    if (!b) return;
    if (batch_check_dev(b, bbx_dev_run_stage(b->dev, &prm, BBX_STAGE_INTEGRATE_V))) return;
    if (batch_download(b, BBX_DL_BODIES)) return;
    zero_host_forces(b);                                             /* :1917-1924 */
}

/* callbacks from a contact list in emission order (same rules as replay_callbacks) */
static void replay_contact_callbacks(BbxWorldPriv* p)
{
    CollisionCollection* cc = &p->collisions;
    SphereSystem* S = &p->spheres; BoxSystem* B = &p->boxes;
    int prev_a = -1, prev_b = -1;
    for (int i = 0; i < cc->count; i++) {
        CollisionContact* k = &cc->contacts[i];
        int a = k->bodyA_index, c = k->bodyB_index;
        if (k->bodyA_type == 1 && k->bodyB_type == 1) {
            if (a != prev_a || c != prev_b) {
                if (B->collision_callbacks[a]) B->collision_callbacks[a](a, 1, c, k);
                if (B->collision_callbacks[c]) B->collision_callbacks[c](c, 1, a, k);
            }
            prev_a = a; prev_b = c;
            continue;
        }
        prev_a = -1; prev_b = -1;
        if (k->bodyA_type == 0 && k->bodyB_type == 0) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 0, c, k);
            if (S->collision_callbacks[c]) S->collision_callbacks[c](c, 0, a, k);
        } else if (k->bodyA_type == 0 && k->bodyB_type == 1) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 1, c, k);
            if (B->collision_callbacks[c]) B->collision_callbacks[c](c, 0, a, k);
        } else if (k->bodyA_type == 0 && k->bodyB_type == 2) {
            if (S->collision_callbacks[a]) S->collision_callbacks[a](a, 2, 0, k);
        }
    }
}

void CollectCollisions(PhysicsWorld* world)                          /* :1979 */
{
    if (!world || !world->collisions) return;
    BbxStepParams prm; struct PhysicsWorldBatch* b = stage_begin(world, &prm);
    if (!b) return;
    world->collisions->count = 0;
    if (batch_check_dev(b, bbx_dev_run_stage(b->dev, &prm, BBX_STAGE_COLLECT))) return;
    int n = 0;
    if (batch_check_dev(b, bbx_dev_download_contacts(b->dev, world->collisions->contacts, world->collisions->capacity, &n))) return;
    world->collisions->count = n;
    if (b->n_callbacks > 0) replay_contact_callbacks(&b->worlds[0]);
}

void GenerateContactConstraints(PhysicsWorld* world)                 /* :1176-1211: a field-by-field copy, no arithmetic */
{
    if (!world || !world->collisions || !world->constraints) return;
    ConstraintCollection* K = world->constraints; CollisionCollection* C = world->collisions;
    K->count = 0;
    for (int i = 0; i < C->count && K->count < K->capacity; i++) {
        CollisionContact* s = &C->contacts[i]; ContactConstraint* c = &K->constraints[K->count++];
        c->bodyA_type = s->bodyA_type; c->bodyA_index = s->bodyA_index; c->bodyB_type = s->bodyB_type; c->bodyB_index = s->bodyB_index;
        c->contact_point = s->contact_point; c->contact_normal = s->contact_normal; c->penetration = s->penetration;
        c->restitution = world->settings.restitution; c->friction = world->settings.friction;
        c->normal_impulse = 0.0f; c->tangent_impulse[0] = 0.0f; c->tangent_impulse[1] = 0.0f;
    }
}

/* the device sweep uses one restitution / friction for the whole list (what GenerateContactConstraints writes) */
static int uniform_materials(struct PhysicsWorldBatch* b, ConstraintCollection* K, float* rest, float* fric)
{
    *rest = b->worlds[0].pub.settings.restitution; *fric = b->worlds[0].pub.settings.friction;
    if (K->count > 0) { *rest = K->constraints[0].restitution; *fric = K->constraints[0].friction; }
    for (int i = 1; i < K->count; i++)
        if (K->constraints[i].restitution != *rest || K->constraints[i].friction != *fric) {
            batch_fail(b, "per-constraint restitution/friction are not supported by the device solver stages");
            return 0;
        }
    return 1;
}

static void solver_stage(PhysicsWorld* world, int stage, int solved, int download_bodies)
{
    if (!world || !world->constraints) return;
    BbxStepParams prm; struct PhysicsWorldBatch* b = stage_begin(world, &prm);
    if (!b) return;
    ConstraintCollection* K = world->constraints;
    float rest, fric;
    if (!uniform_materials(b, K, &rest, &fric)) return;
    prm.settings.restitution = rest; prm.settings.friction = fric;
    if (batch_check_dev(b, bbx_dev_import_list(b->dev, NULL, K->constraints, K->count))) return;
    if (batch_check_dev(b, bbx_dev_run_stage(b->dev, &prm, stage))) return;
    if (batch_check_dev(b, bbx_dev_export_list(b->dev, K->constraints, K->count, rest, fric, solved))) return;
    if (download_bodies) batch_download(b, BBX_DL_BODIES);
}

void PreSolveConstraints(PhysicsWorld* world)      { solver_stage(world, BBX_STAGE_PRESOLVE, 0, 0); }     /* :1295 */
void SolveVelocityConstraints(PhysicsWorld* world) { solver_stage(world, BBX_STAGE_SOLVE_SWEEP, 1, 1); }  /* :1410 */

void SolveConstraints(PhysicsWorld* world)                           /* :1213-1226 */
{
    if (!world || !world->constraints) return;
    GenerateContactConstraints(world);
    solver_stage(world, BBX_STAGE_SOLVE_ALL, world->settings.solver_iterations > 0, 1);
}

static void body_stage(PhysicsWorld* world, int stage)
{
    BbxStepParams prm; struct PhysicsWorldBatch* b = stage_begin(world, &prm);
    if (!b) return;
    if (batch_check_dev(b, bbx_dev_run_stage(b->dev, &prm, stage))) return;
    batch_download(b, BBX_DL_BODIES);
}
void IntegratePositions(PhysicsWorld* world) { body_stage(world, BBX_STAGE_INTEGRATE_P); }               /* :1928 */
void UpdateSleepStates(PhysicsWorld* world)  { body_stage(world, BBX_STAGE_SLEEP); }                      /* :1696 */
void CleanupPhysics(PhysicsWorld* world) { if (world && world->collisions) world->collisions->count = 0; } /* :2305 */

/* ------------------------------------------------------------------------- */
/* public stepping API                                                       */
/* ------------------------------------------------------------------------- */
void PhysicsStep(PhysicsWorld* world, float deltaTime)              /* ballbox.h:1787 */
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return;
    if (p->batch->n_worlds != 1) { batch_fail(p->batch, "PhysicsStep on a member of a batch: use PhysicsStepBatch"); return; }
    batch_step(p->batch, deltaTime, 1);
}

int PhysicsStepN(PhysicsWorld* world, float dt, int steps)
{
    BbxWorldPriv* p = priv_of(world);
    if (!p) return BBX_ERR_BAD_ARGUMENT;
    if (p->batch->n_worlds != 1) { batch_fail(p->batch, "PhysicsStepN on a member of a batch: use PhysicsStepBatch"); return BBX_ERR_BAD_ARGUMENT; }
    return batch_step(p->batch, dt, steps);
}

void BallboxSetSyncMode(PhysicsWorld* w, int mode)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return;
    host_edit_begin(w);                  /* leaving RESIDENT: the host arrays become the truth, so make them current first */
    p->batch->sync_mode = mode; p->batch->uploaded = 0;
}
int  BallboxUpload(PhysicsWorld* w) { BbxWorldPriv* p = priv_of(w); return p ? batch_upload(p->batch) : BBX_ERR_BAD_ARGUMENT; }
int  BallboxDownload(PhysicsWorld* w, int what) { BbxWorldPriv* p = priv_of(w); return p ? batch_download(p->batch, what) : BBX_ERR_BAD_ARGUMENT; }
int  BallboxSynchronize(PhysicsWorld* w)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return BBX_ERR_BAD_ARGUMENT;
    return p->batch->dev ? batch_check_dev(p->batch, bbx_dev_synchronize(p->batch->dev)) : 0;
}
int  BallboxSetDevice(PhysicsWorld* w, int device)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return BBX_ERR_BAD_ARGUMENT;
    if (p->batch->dev) { batch_fail(p->batch, "BallboxSetDevice after the device context was created"); return BBX_ERR_BAD_ARGUMENT; }
    p->batch->device = device;
    return 0;
}
void BallboxSetStream(PhysicsWorld* w, void* stream)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return;
    p->batch->stream = stream;
    if (p->batch->dev) bbx_dev_set_stream(p->batch->dev, stream);
}
const char* BallboxGetLastError(PhysicsWorld* w) { BbxWorldPriv* p = priv_of(w); return (p && p->batch->has_err) ? p->batch->err : NULL; }
void BallboxClearError(PhysicsWorld* w)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return;
    p->batch->has_err = 0; p->batch->err[0] = 0;
    if (p->batch->dev) bbx_dev_clear_error(p->batch->dev);
}
void BallboxSetWarmStarting(PhysicsWorld* w, float factor) { BbxWorldPriv* p = priv_of(w); if (p) p->batch->warm_start = factor > 0.0f ? factor : 0.0f; }
void BallboxSetSolverSchedule(PhysicsWorld* w, int schedule) { BbxWorldPriv* p = priv_of(w); if (p) p->batch->schedule = schedule; }
int  BallboxGetStats(PhysicsWorld* w, BallboxStats* out)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p || !out) return BBX_ERR_BAD_ARGUMENT;
    memset(out, 0, sizeof(*out));
    return p->batch->dev ? batch_check_dev(p->batch, bbx_dev_stats(p->batch->dev, out)) : 0;
}

void BallboxSetStageTiming(PhysicsWorld* w, int on)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p) return;
    p->batch->timing = on;
    if (p->batch->dev) bbx_dev_set_timing(p->batch->dev, on);
}
int BallboxGetStageTiming(PhysicsWorld* w, float* out6, int* steps)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p || !out6) return BBX_ERR_BAD_ARGUMENT;
    for (int k = 0; k < 6; k++) out6[k] = 0.0f;
    if (steps) *steps = 0;
    return p->batch->dev ? batch_check_dev(p->batch, bbx_dev_get_timing(p->batch->dev, out6, steps)) : 0;
}

int BallboxGetReplayTiming(PhysicsWorld* w, float* ms, int* steps)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p || !ms || !p->batch->dev) return BBX_ERR_BAD_ARGUMENT;
    return batch_check_dev(p->batch, bbx_dev_get_replay_timing(p->batch->dev, ms, steps));
}

int BallboxDebugLevels(PhysicsWorld* w, int* counts8, unsigned long long* ns, int max_levels)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p || !p->batch->dev) return BBX_ERR_BAD_ARGUMENT;
    return batch_check_dev(p->batch, bbx_dev_debug_levels(p->batch->dev, counts8, ns, max_levels));
}

int BallboxDebugSchedule(PhysicsWorld* w, int* out4, int cap_nodes)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p || !p->batch->dev) return BBX_ERR_BAD_ARGUMENT;
    return bbx_dev_debug_schedule(p->batch->dev, out4, cap_nodes);
}

int BallboxDebugDeferred(PhysicsWorld* w)
{
    BbxWorldPriv* p = priv_of(w);
    return (p && p->batch->dev) ? bbx_dev_debug_deferred(p->batch->dev) : 0;
}

int BallboxDebugWarpTrace(PhysicsWorld* w, unsigned long long* out, int n_warps)
{
    BbxWorldPriv* p = priv_of(w);
    if (!p || !p->batch->dev) return BBX_ERR_BAD_ARGUMENT;
    return batch_check_dev(p->batch, bbx_dev_debug_warp_trace(p->batch->dev, out, n_warps));
}

static int batch_ok(PhysicsWorldBatch* b) { return b && b->magic == BBX_MAGIC; }
int  BatchSetDevice(PhysicsWorldBatch* b, int device) { return batch_ok(b) ? BallboxSetDevice(&b->worlds[0].pub, device) : BBX_ERR_BAD_ARGUMENT; }
void BatchSetStream(PhysicsWorldBatch* b, void* s) { if (batch_ok(b)) BallboxSetStream(&b->worlds[0].pub, s); }
int  BatchUpload(PhysicsWorldBatch* b) { return batch_ok(b) ? batch_upload(b) : BBX_ERR_BAD_ARGUMENT; }
int  BatchDownload(PhysicsWorldBatch* b, int what) { return batch_ok(b) ? batch_download(b, what) : BBX_ERR_BAD_ARGUMENT; }
int  PhysicsStepBatch(PhysicsWorldBatch* b, float dt, int steps)
{
    if (!batch_ok(b)) return BBX_ERR_BAD_ARGUMENT;
    int keep = b->sync_mode;
    b->sync_mode = BBX_SYNC_RESIDENT;                /* batches are always device resident */
    int rc = batch_step(b, dt, steps);
    b->sync_mode = keep;
    return rc;
}
int  BatchSynchronize(PhysicsWorldBatch* b) { return batch_ok(b) ? BallboxSynchronize(&b->worlds[0].pub) : BBX_ERR_BAD_ARGUMENT; }
int  BatchGetStats(PhysicsWorldBatch* b, BallboxStats* out) { return batch_ok(b) ? BallboxGetStats(&b->worlds[0].pub, out) : BBX_ERR_BAD_ARGUMENT; }
void BatchSetStageTiming(PhysicsWorldBatch* b, int on) { if (batch_ok(b)) BallboxSetStageTiming(&b->worlds[0].pub, on); }
int  BatchGetStageTiming(PhysicsWorldBatch* b, float* out6, int* steps) { return batch_ok(b) ? BallboxGetStageTiming(&b->worlds[0].pub, out6, steps) : BBX_ERR_BAD_ARGUMENT; }
const char* BatchGetLastError(PhysicsWorldBatch* b) { return (batch_ok(b) && b->has_err) ? b->err : NULL; }

int BatchGetDeviceViews(PhysicsWorldBatch* b, BallboxDeviceViews* out)
{
    if (!batch_ok(b) || !out) return BBX_ERR_BAD_ARGUMENT;
    int rc = batch_ensure_dev(b);
    if (rc) return rc;
    BbxDeviceViews v;
    rc = batch_check_dev(b, bbx_dev_views(b->dev, &v));
    if (rc) return rc;
    out->n_worlds = v.n_worlds; out->cap_spheres = v.cap_spheres; out->cap_boxes = v.cap_boxes;
    out->pos = v.pos; out->vel = v.vel; out->box_rot = v.box_rot; out->flags = v.flags;
    out->sphere_force = v.sphere_force; out->sphere_torque = v.sphere_torque;
    out->box_force = v.box_force; out->box_torque = v.box_torque;
    return 0;
}
int BatchAddWrenchDevice(PhysicsWorldBatch* b, const float* sf, const float* st, const float* bf, const float* bt)
{
    if (!batch_ok(b)) return BBX_ERR_BAD_ARGUMENT;
    int rc;
    if (!b->uploaded && (rc = batch_upload(b))) return rc;
    b->dev_ahead = 1;
    return batch_check_dev(b, bbx_dev_add_wrench(b->dev, sf, st, bf, bt));
}
int BatchSnapshot(PhysicsWorldBatch* b)
{
    if (!batch_ok(b)) return BBX_ERR_BAD_ARGUMENT;
    int rc;
    if (!b->uploaded && (rc = batch_upload(b))) return rc;
    return batch_check_dev(b, bbx_dev_snapshot(b->dev));
}
int BatchResetWorlds(PhysicsWorldBatch* b, const int* ids, int n)
{
    if (!batch_ok(b) || !b->dev) return BBX_ERR_BAD_ARGUMENT;
    b->dev_ahead = 1;
    return batch_check_dev(b, bbx_dev_restore_worlds(b->dev, ids, n));
}

const char* BallboxBackendName(void) { return "ballbox-b200/cuda-sm_100a"; }
int BallboxDeviceCount(void) { return bbx_device_count(); }
