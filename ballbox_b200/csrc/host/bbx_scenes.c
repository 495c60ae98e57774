/*
 * bbx_scenes.c -- synthetic scene builders for the five BASELINE.json configs.
 *
 * Uses only the public Ballbox API (AddSphere, AddBox, SetBoxStatic, settings,
 * UseSDF), so the same source builds the same scene on every backend: it is
 * compiled into libballbox_b200.so (the product) and, by oracle/Makefile, into
 * the CPU checkers.  Random numbers come from one LCG:
 *     s = s*1664525 + 1013904223;  u = (s >> 8) / 2^24
 */
#include "ballbox.h"
#include "bbx_sdf_library.h"

#define BBX_SCENE_README   1   /* config 1: one sphere above a static ground box   */
#define BBX_SCENE_BALLS    2   /* config 2: n spheres inside six static walls      */
#define BBX_SCENE_MIXED    3   /* config 3: n/2 spheres + n/2 rotated boxes, walls */
#define BBX_SCENE_SDF      4   /* config 4: n/2 + n/2 on the wavy SDF, four walls  */
#define BBX_SCENE_RLWORLD  5   /* config 5: 128 spheres + 122 boxes + 6 walls      */

typedef struct { unsigned s; } Lcg;
static float lcg_u(Lcg* g)
{
    g->s = g->s * 1664525u + 1013904223u;
    return (float)(g->s >> 8) * (1.0f / 16777216.0f);
}
static float lcg_range(Lcg* g, float lo, float hi) { return lo + (hi - lo) * lcg_u(g); }

/* host-side SDFs handed to UseSDF on CPU backends */
float BbxHostSdfWavyDet(Vec3 p)
{
    float prm[1] = {0};
    return bbx_sdf_eval(BBX_SDF_ID_WAVY_DET, prm, p.x, p.y, p.z);
}
float BbxHostSdfWavyLibm(Vec3 p)
{
    float prm[1] = {0};
    return bbx_sdf_eval(BBX_SDF_ID_WAVY_LIBM, prm, p.x, p.y, p.z);
}
float BbxHostSdfPlane0(Vec3 p) { return p.y; }

void BbxSceneUseSdf(PhysicsWorld* w, int sdf_id)
{
    float prm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (sdf_id == BBX_SDF_ID_WAVY_DET)       UseSDF(w, BbxHostSdfWavyDet);
    else if (sdf_id == BBX_SDF_ID_WAVY_LIBM) UseSDF(w, BbxHostSdfWavyLibm);
    else if (sdf_id == BBX_SDF_ID_PLANE_Y)   UseSDF(w, BbxHostSdfPlane0);
    else { UseSDF(w, NULL); return; }
#ifdef BBX_PRODUCT
    UseSDFDevice(w, sdf_id, prm);
#else
    (void)prm;
#endif
}

static Quat random_rotation(Lcg* g)
{
    Vec3 axis = { lcg_range(g, -1.0f, 1.0f), lcg_range(g, -1.0f, 1.0f), lcg_range(g, -1.0f, 1.0f) };
    if (axis.x * axis.x + axis.y * axis.y + axis.z * axis.z < 1e-4f) { axis.x = 0; axis.y = 1; axis.z = 0; }
    float angle = lcg_u(g) * PHYSICS_PI;
    return QuatFromAxisAngle(axis, angle);
}

static void add_static_box(PhysicsWorld* w, Vec3 c, Vec3 h)
{
    int b = AddBox(w, c, h, QuatIdentity());
    SetBoxStatic(w, b, true);
}

/* six (or four) static walls of half-thickness 1 around a cube of half-side L;
 * neighbouring walls overlap at the edges, which produces the constant set of
 * static-static contacts the reference emits (ballbox.h:2011-2047 has no filter). */
static void add_walls(PhysicsWorld* w, float L, int with_floor_and_ceiling, float ymid, float yhalf)
{
    float e = L + 2.0f;
    if (with_floor_and_ceiling) {
        add_static_box(w, (Vec3){0, -L - 1.0f, 0}, (Vec3){e, 1.0f, e});
        add_static_box(w, (Vec3){0,  L + 1.0f, 0}, (Vec3){e, 1.0f, e});
        ymid = 0.0f; yhalf = e;
    }
    add_static_box(w, (Vec3){-L - 1.0f, ymid, 0}, (Vec3){1.0f, yhalf, e});
    add_static_box(w, (Vec3){ L + 1.0f, ymid, 0}, (Vec3){1.0f, yhalf, e});
    add_static_box(w, (Vec3){0, ymid, -L - 1.0f}, (Vec3){e, yhalf, 1.0f});
    add_static_box(w, (Vec3){0, ymid,  L + 1.0f}, (Vec3){e, yhalf, 1.0f});
}

/* capacities a world needs for scene `kind` with n dynamic bodies */
void BbxSceneCapacity(int kind, int n, int* max_spheres, int* max_boxes, int* max_collisions)
{
    int s = 0, b = 0, c = 0;
    switch (kind) {
    case BBX_SCENE_README:  s = 10; b = 10; c = 50; break;
    case BBX_SCENE_BALLS:   s = n; b = 6; c = 8 * n + 4096; break;
    case BBX_SCENE_MIXED:   s = n / 2; b = n - n / 2 + 6; c = 12 * n + 4096; break;
    case BBX_SCENE_SDF:     s = n / 2; b = n - n / 2 + 4; c = 12 * n + 4096; break;
    case BBX_SCENE_RLWORLD: s = 128; b = 128; c = 4096; break;
    default: break;
    }
    if (max_spheres) *max_spheres = s;
    if (max_boxes) *max_boxes = b;
    if (max_collisions) *max_collisions = c;
}

/* Populate an empty world.  Returns the number of bodies added, -1 on error. */
int BbxSceneBuild(PhysicsWorld* w, int kind, int n, unsigned seed)
{
    if (!w) return -1;
    Lcg g = { seed };
    w->gravity = (Vec3){0.0f, -9.8f, 0.0f};

    if (kind == BBX_SCENE_README) {                      /* reference README.md:43-55 */
        AddSphere(w, (Vec3){0, 5, 0}, 1.0f);
        int ground = AddBox(w, (Vec3){0, -1, 0}, (Vec3){10, 1, 10}, QuatIdentity());
        SetBoxStatic(w, ground, true);
        return 2;
    }

    if (kind == BBX_SCENE_RLWORLD) n = 250;
    float L = 0.8f * cbrtf((float)n);
    int n_boxes = (kind == BBX_SCENE_BALLS) ? 0 : (kind == BBX_SCENE_RLWORLD ? 122 : n - n / 2);
    int n_spheres = n - n_boxes;

    w->settings.restitution = 0.3f;
    w->settings.solver_iterations = 8;
    if (kind != BBX_SCENE_BALLS) w->settings.friction = 0.3f;

    float ylo = -(L - 0.75f), yhi = L - 0.75f;
    if (kind == BBX_SCENE_SDF) {
        ylo = 1.5f; yhi = 1.5f + 2.0f * L;
        add_walls(w, L, 0, L, L + 4.0f);
        BbxSceneUseSdf(w, BBX_SDF_ID_WAVY_DET);
    } else {
        add_walls(w, L, 1, 0, 0);
    }

    float lim = L - 0.75f;
    int added = 0, si = 0, bi = 0;
    while (si < n_spheres || bi < n_boxes) {
        /* interleave so that sphere and box indices are both spatially random */
        int want_box = (bi < n_boxes) && (si >= n_spheres || ((si + bi) & 1));
        Vec3 p = { lcg_range(&g, -lim, lim), lcg_range(&g, ylo, yhi), lcg_range(&g, -lim, lim) };
        if (want_box) {
            Quat q = random_rotation(&g);
            if (AddBox(w, p, (Vec3){0.4f, 0.4f, 0.4f}, q) < 0) return -1;
            bi++;
        } else {
            if (AddSphere(w, p, 0.5f) < 0) return -1;
            si++;
        }
        added++;
    }
    return added + ((kind == BBX_SCENE_SDF) ? 4 : 6);
}

/* FNV-1a over the full mutable state; used to compare backends bit-for-bit. */
unsigned long long BbxStateHash(PhysicsWorld* w)
{
    unsigned long long h = 1469598103934665603ull;
#define MIX(ptr, bytes) do { const unsigned char* q_ = (const unsigned char*)(ptr); \
        for (size_t i_ = 0; i_ < (size_t)(bytes); i_++) { h ^= q_[i_]; h *= 1099511628211ull; } } while (0)
    SphereSystem* s = w->spheres; BoxSystem* b = w->boxes;
    MIX(s->positions, sizeof(Vec3) * s->count);
    MIX(s->velocities, sizeof(Vec3) * s->count);
    MIX(s->angular_velocities, sizeof(Vec3) * s->count);
    MIX(s->is_sleeping, s->count);
    MIX(s->sleep_timers, sizeof(float) * s->count);
    MIX(b->positions, sizeof(Vec3) * b->count);
    MIX(b->rotations, sizeof(Quat) * b->count);
    MIX(b->velocities, sizeof(Vec3) * b->count);
    MIX(b->angular_velocities, sizeof(Vec3) * b->count);
    MIX(b->is_sleeping, b->count);
    MIX(b->sleep_timers, sizeof(float) * b->count);
#undef MIX
    return h;
}
