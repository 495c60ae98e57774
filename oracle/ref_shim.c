/*
 * ref_shim.c -- TEST INFRASTRUCTURE.  Builds the UNMODIFIED reference header
 * (found through -I$(BALLBOX_REF_DIR), never copied into this repo) into a
 * shared library, oracle/_ref/libballbox_ref.so, together with
 *   - the shared scene builders (bbx_scenes.c, public API only), and
 *   - RefPhysicsStepPruned: the reference's own PhysicsStep stage sequence
 *     (ballbox.h:1797-1816) in which ONLY the loop bounds of CollectCollisions
 *     (ballbox.h:1986-1987, :2011-2012, :2050-2051) are replaced by a CPU grid
 *     that skips pairs whose bounding spheres cannot touch.  Narrowphase,
 *     manifold, callbacks, presolve and solver are the reference's functions.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load the result.
 */
#define BALLBOX_IMPLEMENTATION
#include "ballbox.h"                                  /* the reference's, via -I */
#include "../ballbox_b200/csrc/host/bbx_scenes.c"
#include "prune_grid.h"

static float ref_box_bound(const BoxSystem* b, int i)
{
    Quat q = b->rotations[i];
    float q2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (q2 < 1.0f) q2 = 1.0f;
    return Vec3Length(b->sizes[i]) * q2 * 1.0005f + 1e-5f;
}

/* radius above which a body bypasses the grid: the largest non-static body */
static float ref_small_cut(PhysicsWorld* w, const float* brad)
{
    float cut = 0.0f; int any = 0;
    for (int i = 0; i < w->spheres->count; i++) if (!w->spheres->is_static[i]) { any = 1; if (w->spheres->radii[i] > cut) cut = w->spheres->radii[i]; }
    for (int i = 0; i < w->boxes->count; i++) if (!w->boxes->is_static[i]) { any = 1; if (brad[i] > cut) cut = brad[i]; }
    if (!any) {
        for (int i = 0; i < w->spheres->count; i++) if (w->spheres->radii[i] > cut) cut = w->spheres->radii[i];
        for (int i = 0; i < w->boxes->count; i++) if (brad[i] > cut) cut = brad[i];
    }
    return cut > 0.0f ? cut : 1.0f;
}

long long g_ref_pairs_tested = 0;

void RefCollectCollisionsPruned(PhysicsWorld* world)
{
    if (!world || !world->collisions) return;
    SphereSystem* S = world->spheres; BoxSystem* B = world->boxes;
    CollisionCollection* C = world->collisions;
    C->count = 0;

    float* brad = (float*)malloc(sizeof(float) * (B->count > 0 ? B->count : 1));
    for (int i = 0; i < B->count; i++) brad[i] = ref_box_bound(B, i);
    float cut = ref_small_cut(world, brad);
    PruneGrid* gs = pg_build(S->count, (const float*)S->positions, S->radii, cut);
    PruneGrid* gb = pg_build(B->count, (const float*)B->positions, brad, cut);
    int* cand = NULL; int cap = 0;

    /* sphere-sphere, same body as reference :1986-2008 */
    for (int i = 0; i < S->count; i++) {
        int nc = pg_query(gs, (const float*)&S->positions[i], S->radii[i], i + 1, &cand, &cap);
        for (int t = 0; t < nc; t++) {
            int j = cand[t];
            if (C->count >= C->capacity) break;
            g_ref_pairs_tested++;
            CollisionContact* contact = &C->contacts[C->count];
            if (CheckSphereCollisionWithData(S->positions[i], S->radii[i], S->positions[j], S->radii[j], contact)) {
                contact->bodyA_type = 0; contact->bodyA_index = i;
                contact->bodyB_type = 0; contact->bodyB_index = j;
                C->count++;
                if (S->collision_callbacks[i]) S->collision_callbacks[i](i, 0, j, contact);
                if (S->collision_callbacks[j]) S->collision_callbacks[j](j, 0, i, contact);
            }
        }
    }
    /* box-box manifolds, reference :2011-2047 */
    for (int i = 0; i < B->count; i++) {
        int nc = pg_query(gb, (const float*)&B->positions[i], brad[i], i + 1, &cand, &cap);
        for (int t = 0; t < nc; t++) {
            int j = cand[t];
            g_ref_pairs_tested++;
            ContactManifold manifold;
            GenerateBoxBoxManifold(world, i, j, &manifold);
            if (manifold.point_count > 0) {
                for (int k = 0; k < manifold.point_count; k++) {
                    if (C->count >= C->capacity) break;
                    CollisionContact* contact = &C->contacts[C->count];
                    contact->bodyA_type = manifold.bodyA_type; contact->bodyA_index = manifold.bodyA_index;
                    contact->bodyB_type = manifold.bodyB_type; contact->bodyB_index = manifold.bodyB_index;
                    contact->contact_point = manifold.points[k];
                    contact->contact_normal = manifold.normal;
                    contact->penetration = manifold.penetrations[k];
                    C->count++;
                }
                if (B->collision_callbacks[i]) B->collision_callbacks[i](i, 1, j, &C->contacts[C->count - manifold.point_count]);
                if (B->collision_callbacks[j]) B->collision_callbacks[j](j, 1, i, &C->contacts[C->count - manifold.point_count]);
            }
        }
    }
    /* sphere-box, reference :2050-2072 */
    for (int i = 0; i < S->count; i++) {
        int nc = pg_query(gb, (const float*)&S->positions[i], S->radii[i], 0, &cand, &cap);
        for (int t = 0; t < nc; t++) {
            int j = cand[t];
            if (C->count >= C->capacity) break;
            g_ref_pairs_tested++;
            CollisionContact* contact = &C->contacts[C->count];
            if (CheckSphereBoxCollisionWithData(S->positions[i], S->radii[i], B->positions[j], B->sizes[j], B->rotations[j], contact)) {
                contact->bodyA_type = 0; contact->bodyA_index = i;
                contact->bodyB_type = 1; contact->bodyB_index = j;
                C->count++;
                if (S->collision_callbacks[i]) S->collision_callbacks[i](i, 1, j, contact);
                if (B->collision_callbacks[j]) B->collision_callbacks[j](j, 0, i, contact);
            }
        }
    }
    free(cand); pg_free(gs); pg_free(gb); free(brad);

    /* SDF phases are per body, nothing to prune: reference :2074-2114 verbatim behaviour */
    if (world->usesSDF && world->SDFCollider) {
        for (int i = 0; i < S->count; i++) {
            if (C->count >= C->capacity) break;
            Vec3 c = S->positions[i]; float r = S->radii[i];
            float d = world->SDFCollider(c);
            if (d < r) {
                CollisionContact* contact = &C->contacts[C->count];
                contact->penetration = r - d;
                Vec3 grad = SDF_EstimateNormal(world, c);
                contact->contact_normal = Vec3Scale(grad, -1.0f);
                contact->contact_point = Vec3Sub(c, Vec3Scale(contact->contact_normal, r));
                contact->bodyA_type = 0; contact->bodyA_index = i;
                contact->bodyB_type = 2; contact->bodyB_index = 0;
                C->count++;
                if (S->collision_callbacks[i]) S->collision_callbacks[i](i, 2, 0, contact);
            }
        }
        for (int i = 0; i < B->count; i++) GenerateBoxSDFContacts(world, i);
    }
}

/* the reference's PhysicsStep (ballbox.h:1787-1817) with the pruned collect */
void RefPhysicsStepPruned(PhysicsWorld* world, float deltaTime)
{
    if (!world) return;
    if (deltaTime <= 0.0f || deltaTime < 1e-6f) return;
    world->dt = deltaTime;
    ApplyForces(world);
    IntegrateVelocities(world);
    RefCollectCollisionsPruned(world);
    SolveConstraints(world);
    IntegratePositions(world);
    UpdateSleepStates(world);
    CleanupPhysics(world);
}

/* detection only (pruned or full), leaving world->collisions filled */
void RefDetectOnly(PhysicsWorld* world, int pruned)
{
    if (pruned) RefCollectCollisionsPruned(world); else CollectCollisions(world);
}

const char* BallboxBackendName(void) { return "reference/unmodified-header"; }

/* struct sizes for the ABI test */
int RefSizeof(int which)
{
    switch (which) {
    case 0: return (int)sizeof(Vec3);             case 1: return (int)sizeof(Quat);
    case 2: return (int)sizeof(CollisionContact); case 3: return (int)sizeof(ContactConstraint);
    case 4: return (int)sizeof(ContactManifold);  case 5: return (int)sizeof(PhysicsSettings);
    case 6: return (int)sizeof(SphereSystem);     case 7: return (int)sizeof(BoxSystem);
    case 8: return (int)sizeof(PhysicsWorld);
    default: return -1;
    }
}
