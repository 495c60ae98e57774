// bbx_queries.cu -- the standalone queries and helpers of the public API that are not part of
// PhysicsStep (reference ballbox.h:231-239, :262-297).  They are single-pair / single-body host
// functions in the reference and stay host functions here; the geometry is NOT a second
// implementation: it is bbx_geom.cuh, the same __host__ __device__ code the narrowphase kernels
// run, compiled for the host without FMA contraction, so a query returns the bits the step
// would produce for that pair.  Functions that walk the user's SDF call the host pointer
// (world->SDFCollider), exactly like the reference.
//
//   Check*Collision(+WithData), Check*SystemCollisions   :447-491, :1114-1132, :1539-1626
//   GenerateBoxBoxManifold, SelectOptimalContactPoints,
//   AddManifoldToConstraints                              :2550-2869
//   SDF_EstimateNormal, CheckBoxSDFCollision,
//   GenerateBoxSDFContacts                                :2318-2477
//   GetBodyVelocityAtPoint, CalculateEffectiveMass,
//   ApplyImpulse (legacy impulse path), SolvePositionConstraints   :2122-2297, :1491-1535
// ResolveCollisions and PositionalCorrection are declared but never defined by the reference
// (:221, :239) and are not provided.
#include <stdlib.h>
#include <vector>
#include "bbx_dev.cuh"
#include "bbx_geom.cuh"

static inline float3 F3(Vec3 v) { return make_float3(v.x, v.y, v.z); }
static inline float4 F4(Quat q) { return make_float4(q.x, q.y, q.z, q.w); }
static inline Vec3 V3(float3 v) { Vec3 r = { v.x, v.y, v.z }; return r; }

extern "C" {

// ---- sphere / box pair queries ---------------------------------------------------------------
bool CheckSphereCollisionWithData(Vec3 posA, float radiusA, Vec3 posB, float radiusB, CollisionContact* contact)
{
    float3 pt, nrm; float pen;
    if (!sphere_sphere_contact(F3(posA), radiusA, F3(posB), radiusB, pt, nrm, pen)) return false;
    if (contact) { contact->contact_normal = V3(nrm); contact->contact_point = V3(pt); contact->penetration = pen; }
    return true;
}
bool CheckSphereCollision(Vec3 p1, float r1, Vec3 p2, float r2) { return CheckSphereCollisionWithData(p1, r1, p2, r2, NULL); }

bool CheckSphereSystemCollisions(SphereSystem* system, int sphere_index)
{
    if (!system || sphere_index < 0 || sphere_index >= system->count) return false;
    for (int i = 0; i < system->count; i++)
        if (i != sphere_index && CheckSphereCollision(system->positions[sphere_index], system->radii[sphere_index],
                                                      system->positions[i], system->radii[i])) return true;
    return false;
}

bool CheckSphereBoxCollisionWithData(Vec3 spherePos, float sphereRadius, Vec3 boxPos, Vec3 boxSize, Quat boxRot, CollisionContact* contact)
{
    float3 pt, nrm; float pen;
    if (!sphere_box_contact(F3(spherePos), sphereRadius, F3(boxPos), F3(boxSize), F4(boxRot), pt, nrm, pen)) return false;
    if (contact) { contact->contact_normal = V3(nrm); contact->contact_point = V3(pt); contact->penetration = pen; }
    return true;
}
bool CheckSphereBoxCollision(Vec3 sp, float sr, Vec3 bp, Vec3 bs, Quat br) { return CheckSphereBoxCollisionWithData(sp, sr, bp, bs, br, NULL); }

// SAT + contact point + A->B normal (:1040-1132); `axisType` and the unflipped axis feed the manifold
static bool box_pair(Vec3 posA, Vec3 sizeA, Quat rotA, Vec3 posB, Vec3 sizeB, Quat rotB,
                     float3& nrm, float3& pt, float& pen, int& axisType, float3& rawAxis)
{
    Obb A = make_obb(F3(posA), F3(sizeA), F4(rotA)), B = make_obb(F3(posB), F3(sizeB), F4(rotB));
    float minPd;
    if (!sat_axes(A, B, minPd, rawAxis, axisType)) return false;
    nrm = rawAxis;
    sat_point(A, B, axisType, nrm, pt);
    pen = fabsf(minPd);
    return true;
}

bool CheckBoxCollisionWithData(Vec3 posA, Vec3 sizeA, Quat rotA, Vec3 posB, Vec3 sizeB, Quat rotB, CollisionContact* contact)
{
    float3 nrm, pt, raw; float pen; int axisType;
    if (!box_pair(posA, sizeA, rotA, posB, sizeB, rotB, nrm, pt, pen, axisType, raw)) return false;
    if (contact) { contact->contact_normal = V3(nrm); contact->contact_point = V3(pt); contact->penetration = pen; }
    return true;
}
bool CheckBoxCollision(Vec3 p1, Vec3 s1, Quat r1, Vec3 p2, Vec3 s2, Quat r2) { return CheckBoxCollisionWithData(p1, s1, r1, p2, s2, r2, NULL); }

bool CheckBoxSystemCollisions(BoxSystem* system, int box_index)
{
    if (!system || box_index < 0 || box_index >= system->count) return false;
    for (int i = 0; i < system->count; i++)
        if (i != box_index && CheckBoxCollision(system->positions[box_index], system->sizes[box_index], system->rotations[box_index],
                                                system->positions[i], system->sizes[i], system->rotations[i])) return true;
    return false;
}

// ---- manifolds ---------------------------------------------------------------------------------
void GenerateBoxBoxManifold(PhysicsWorld* world, int boxIndexA, int boxIndexB, ContactManifold* manifold)
{
    if (!world || !manifold) return;
    BoxSystem* B = world->boxes;
    if (boxIndexA < 0 || boxIndexA >= B->count || boxIndexB < 0 || boxIndexB >= B->count || boxIndexA == boxIndexB) return;
    manifold->point_count = 0;
    manifold->bodyA_type = 1; manifold->bodyA_index = boxIndexA;
    manifold->bodyB_type = 1; manifold->bodyB_index = boxIndexB;
    float3 nrm, pt, raw; float pen; int axisType;
    if (!box_pair(B->positions[boxIndexA], B->sizes[boxIndexA], B->rotations[boxIndexA],
                  B->positions[boxIndexB], B->sizes[boxIndexB], B->rotations[boxIndexB], nrm, pt, pen, axisType, raw)) return;
    int mc = world->settings.manifold_max_contacts;
    if (mc > MAX_MANIFOLD_CONTACTS) mc = MAX_MANIFOLD_CONTACTS;      // the struct holds 8 points (:104-116); the reference would overrun it
    if (mc < 1) mc = 1;
    float3 opt[8]; float open_[8];
    float3 normal = raw;
    int np = box_box_manifold(F3(B->positions[boxIndexA]), F3(B->sizes[boxIndexA]), F4(B->rotations[boxIndexA]),
                              F3(B->positions[boxIndexB]), F3(B->sizes[boxIndexB]), F4(B->rotations[boxIndexB]),
                              world->settings.manifold_contact_tolerance, world->settings.manifold_penetration_tolerance,
                              mc, pen, axisType, opt, open_, normal);
    manifold->normal = V3(normal);
    for (int i = 0; i < np; i++) { manifold->points[i] = V3(opt[i]); manifold->penetrations[i] = open_[i]; }
    manifold->point_count = np;
}

// :2726-2843.  One or two candidates are taken as they are; up to max_contacts: deepest first, the rest in order;
// more: the deepest plus the (max_contacts - 1) farthest from the centroid (ties keep the first seen).
void SelectOptimalContactPoints(Vec3* candidate_points, float* penetrations, int candidate_count,
                                ContactManifold* manifold, Vec3 normal, int max_contacts)
{
    (void)normal;                 // the reference derives two tangents from it and never uses them
    if (!candidate_points || !penetrations || !manifold || candidate_count <= 0) return;
    const int room = MAX_MANIFOLD_CONTACTS;                          // never write past the struct
    if (candidate_count <= 2) {
        for (int i = 0; i < candidate_count; i++) { manifold->points[i] = candidate_points[i]; manifold->penetrations[i] = penetrations[i]; }
        manifold->point_count = candidate_count;
        return;
    }
    int deep = 0;
    for (int i = 1; i < candidate_count; i++) if (penetrations[i] > penetrations[deep]) deep = i;
    manifold->points[0] = candidate_points[deep]; manifold->penetrations[0] = penetrations[deep];
    manifold->point_count = 1;
    if (candidate_count <= max_contacts) {
        int k = 1;
        for (int i = 0; i < candidate_count && k < room; i++)
            if (i != deep) { manifold->points[k] = candidate_points[i]; manifold->penetrations[k] = penetrations[i]; k++; }
        manifold->point_count = k;
        return;
    }
    float3 centre = f3(0.0f, 0.0f, 0.0f);
    for (int i = 0; i < candidate_count; i++) centre = v3add(centre, F3(candidate_points[i]));
    centre = v3scale(centre, 1.0f / (float)candidate_count);
    int rem = max_contacts - 1;
    if (rem <= 0) return;
    std::vector<float> bestd((size_t)rem, -1.0f);
    std::vector<int> besti((size_t)rem, -1);
    for (int i = 0; i < candidate_count; i++) {
        if (i == deep) continue;
        float dist = v3len(v3sub(F3(candidate_points[i]), centre));
        for (int j = 0; j < rem; j++) {
            if (dist > bestd[j]) {
                for (int k = rem - 1; k > j; k--) { bestd[k] = bestd[k - 1]; besti[k] = besti[k - 1]; }
                bestd[j] = dist; besti[j] = i;
                break;
            }
        }
    }
    for (int j = 0; j < rem && manifold->point_count < room; j++)
        if (besti[j] >= 0) {
            manifold->points[manifold->point_count] = candidate_points[besti[j]];
            manifold->penetrations[manifold->point_count] = penetrations[besti[j]];
            manifold->point_count++;
        }
}

void AddManifoldToConstraints(PhysicsWorld* world, ContactManifold* manifold)          // :2845-2869
{
    if (!world || !manifold || manifold->point_count == 0) return;
    ConstraintCollection* cc = world->constraints;
    for (int i = 0; i < manifold->point_count && cc->count < cc->capacity; i++) {
        ContactConstraint* k = &cc->constraints[cc->count++];
        k->bodyA_type = manifold->bodyA_type; k->bodyA_index = manifold->bodyA_index;
        k->bodyB_type = manifold->bodyB_type; k->bodyB_index = manifold->bodyB_index;
        k->contact_point = manifold->points[i];
        k->contact_normal = manifold->normal;
        k->penetration = manifold->penetrations[i];
        k->restitution = world->settings.restitution;
        k->friction = world->settings.friction;
    }
}

// ---- SDF helpers: the user's host function is called directly, as in the reference ------------
static inline float sdf_host(PhysicsWorld* w, float3 p) { return w->SDFCollider(V3(p)); }

Vec3 SDF_EstimateNormal(PhysicsWorld* world, Vec3 point)                                 // :2318-2331
{
    float e = world->settings.sdf_normal_epsilon;
    float3 p = F3(point), ex = f3(e, 0.0f, 0.0f), ey = f3(0.0f, e, 0.0f), ez = f3(0.0f, 0.0f, e);
    float nx = (sdf_host(world, v3add(p, ex)) - sdf_host(world, v3sub(p, ex))) / (2.0f * e);
    float ny = (sdf_host(world, v3add(p, ey)) - sdf_host(world, v3sub(p, ey))) / (2.0f * e);
    float nz = (sdf_host(world, v3add(p, ez)) - sdf_host(world, v3sub(p, ez))) / (2.0f * e);
    return V3(v3norm(f3(nx, ny, nz)));
}

bool CheckBoxSDFCollision(PhysicsWorld* world, Vec3 boxPos, Vec3 boxSize, Quat boxRot, CollisionContact* contact)   // :2334-2410
{
    // circumscribed sphere, then the 8 vertices, then the 6 face centres; the deepest probe wins
    if (sdf_host(world, F3(boxPos)) > v3len(F3(boxSize))) return false;
    Obb o = make_obb(F3(boxPos), F3(boxSize), F4(boxRot));
    float3 probes[14];
    obb_vertices(o, probes);
    float3 c = F3(boxPos); float4 q = F4(boxRot);
    float3 ax = qrot(q, f3(1.0f, 0.0f, 0.0f)), ay = qrot(q, f3(0.0f, 1.0f, 0.0f)), az = qrot(q, f3(0.0f, 0.0f, 1.0f));
    probes[8]  = v3add(c, v3scale(ax, boxSize.x)); probes[9]  = v3sub(c, v3scale(ax, boxSize.x));
    probes[10] = v3add(c, v3scale(ay, boxSize.y)); probes[11] = v3sub(c, v3scale(ay, boxSize.y));
    probes[12] = v3add(c, v3scale(az, boxSize.z)); probes[13] = v3sub(c, v3scale(az, boxSize.z));
    float deepest = -999999.0f; float3 at = f3(0.0f, 0.0f, 0.0f); bool found = false;
    for (int i = 0; i < 14; i++) {
        float d = sdf_host(world, probes[i]);
        if (d < 0.0f) { found = true; if (-d > deepest) { deepest = -d; at = probes[i]; } }
    }
    if (!found || !contact) return false;          // (the reference also answers false when no contact record is given)
    contact->penetration = deepest;
    float3 g = F3(SDF_EstimateNormal(world, V3(at)));
    contact->contact_normal = V3(v3scale(g, -1.0f));
    contact->contact_point = V3(at);
    return true;
}

void GenerateBoxSDFContacts(PhysicsWorld* world, int boxIndex)                          // :2413-2477
{
    if (!world || !world->usesSDF || !world->SDFCollider) return;
    if (boxIndex < 0 || boxIndex >= world->boxes->count) return;
    static const int edge_pairs[12][2] = { {0,1},{1,2},{2,3},{3,0}, {4,5},{5,6},{6,7},{7,4}, {0,4},{1,5},{2,6},{3,7} };
    Obb o = make_obb(F3(world->boxes->positions[boxIndex]), F3(world->boxes->sizes[boxIndex]), F4(world->boxes->rotations[boxIndex]));
    float3 v[8];
    obb_vertices(o, v);
    CollisionCollection* cc = world->collisions;
    for (int k = 0; k < 20; k++) {
        // the reference leaves each of its two loops at capacity; nothing is added after that either way
        if (cc->count >= cc->capacity) break;
        float3 p = (k < 8) ? v[k] : v3scale(v3add(v[edge_pairs[k - 8][0]], v[edge_pairs[k - 8][1]]), 0.5f);
        float d = sdf_host(world, p);
        if (d < -0.001f) {
            CollisionContact* c = &cc->contacts[cc->count];
            c->penetration = -d;
            c->contact_normal = V3(v3scale(F3(SDF_EstimateNormal(world, V3(p))), -1.0f));
            c->contact_point = V3(p);
            c->bodyA_type = 1; c->bodyA_index = boxIndex; c->bodyB_type = 2; c->bodyB_index = 0;
            cc->count++;
        }
    }
}

// ---- legacy impulse path (never called by PhysicsStep) ------------------------------------------
static bool body_fields(PhysicsWorld* w, int type, int index, Vec3** pos, Vec3** vel, Vec3** angvel)
{
    if (type == 0) { *pos = &w->spheres->positions[index]; *vel = &w->spheres->velocities[index]; *angvel = &w->spheres->angular_velocities[index]; return true; }
    if (type == 1) { *pos = &w->boxes->positions[index]; *vel = &w->boxes->velocities[index]; *angvel = &w->boxes->angular_velocities[index]; return true; }
    return false;
}

Vec3 GetBodyVelocityAtPoint(PhysicsWorld* world, int bodyType, int bodyIndex, Vec3 point)     // :2122-2150
{
    Vec3 *pos, *vel, *ang;
    float3 c = f3(0.0f, 0.0f, 0.0f), v = c, w = c;                  // the SDF (type 2) is static
    if (body_fields(world, bodyType, bodyIndex, &pos, &vel, &ang)) { c = F3(*pos); v = F3(*vel); w = F3(*ang); }
    return V3(v3add(v, v3cross(w, v3sub(F3(point), c))));
}

float CalculateEffectiveMass(PhysicsWorld* world, CollisionContact* contact)            // :2153-2238
{
    float3 n = F3(contact->contact_normal);
    float sum = 0.0f, inertia_terms[2] = { 0.0f, 0.0f }, inv_m[2];
    const int types[2] = { contact->bodyA_type, contact->bodyB_type }, idx[2] = { contact->bodyA_index, contact->bodyB_index };
    for (int s = 0; s < 2; s++) {
        const int t = types[s], i = idx[s];
        float3 c = t == 0 ? F3(world->spheres->positions[i]) : t == 1 ? F3(world->boxes->positions[i]) : f3(0.0f, 0.0f, 0.0f);
        bool is_static = t == 0 ? world->spheres->is_static[i] : t == 1 ? world->boxes->is_static[i] : true;
        float3 r = v3sub(F3(contact->contact_point), c);
        inv_m[s] = is_static ? 0.0f : 1.0f / (t == 0 ? world->spheres->masses[i] : world->boxes->masses[i]);
        if (is_static) continue;
        float3 rn = v3cross(r, n);
        if (t == 0) {
            inertia_terms[s] = v3dot(rn, rn) / world->spheres->inertias[i];
        } else {                       // r x n in the box frame, diagonal inertia there
            float3 l = qrot(qconj(F4(world->boxes->rotations[i])), rn);
            Vec3 I = world->boxes->inertias[i];
            inertia_terms[s] = (l.x * l.x) / I.x + (l.y * l.y) / I.y + (l.z * l.z) / I.z;
        }
    }
    sum = inv_m[0] + inv_m[1] + inertia_terms[0] + inertia_terms[1];
    return 1.0f / sum;
}

void ApplyImpulse(PhysicsWorld* world, int bodyType, int bodyIndex, Vec3 impulse, Vec3 contactPoint)   // :2241-2297
{
    if (bodyType == 0) {
        SphereSystem* S = world->spheres; const int i = bodyIndex;
        if (S->is_static[i]) return;
        if (S->is_sleeping[i]) { S->is_sleeping[i] = false; S->sleep_timers[i] = 0.0f; }
        float3 J = F3(impulse);
        S->velocities[i] = V3(v3add(F3(S->velocities[i]), v3scale(J, 1.0f / S->masses[i])));
        float3 torque = v3cross(v3sub(F3(contactPoint), F3(S->positions[i])), J);
        S->angular_velocities[i] = V3(v3add(F3(S->angular_velocities[i]), v3scale(torque, 1.0f / S->inertias[i])));
    } else if (bodyType == 1) {
        BoxSystem* B = world->boxes; const int i = bodyIndex;
        if (B->is_static[i]) return;
        if (B->is_sleeping[i]) { B->is_sleeping[i] = false; B->sleep_timers[i] = 0.0f; }
        float3 J = F3(impulse);
        B->velocities[i] = V3(v3add(F3(B->velocities[i]), v3scale(J, 1.0f / B->masses[i])));
        float3 torque = v3cross(v3sub(F3(contactPoint), F3(B->positions[i])), J);
        float4 q = F4(B->rotations[i]);
        float3 tl = qrot(qconj(q), torque);
        Vec3 I = B->inertias[i];
        float3 dl = f3(tl.x / I.x, tl.y / I.y, tl.z / I.z);
        B->angular_velocities[i] = V3(v3add(F3(B->angular_velocities[i]), qrot(q, dl)));
    }
}

// 1/m of a constraint body: 0 for static bodies and for the SDF (:1264-1277)
static float inv_mass_of(PhysicsWorld* w, int type, int index)
{
    if (type == 0) return (index < 0 || index >= w->spheres->count || w->spheres->is_static[index]) ? 0.0f : 1.0f / w->spheres->masses[index];
    if (type == 1) return (index < 0 || index >= w->boxes->count || w->boxes->is_static[index]) ? 0.0f : 1.0f / w->boxes->masses[index];
    return 0.0f;
}

void SolvePositionConstraints(PhysicsWorld* world)                                      // :1491-1535
{
    if (!world || !world->constraints) return;
    const float allowed = world->settings.allowed_penetration;
    for (int i = 0; i < world->constraints->count; i++) {
        ContactConstraint* c = &world->constraints->constraints[i];
        if (c->penetration <= allowed) continue;
        float imA = inv_mass_of(world, c->bodyA_type, c->bodyA_index), imB = inv_mass_of(world, c->bodyB_type, c->bodyB_index);
        float total = imA + imB;
        if (total < 1e-10f) continue;
        float amount = (c->penetration - allowed) * 0.8f;
        float3 corr = v3scale(F3(c->contact_normal), amount / total);
        if (imA > 0.0f) {
            Vec3* p = c->bodyA_type == 0 ? &world->spheres->positions[c->bodyA_index] : c->bodyA_type == 1 ? &world->boxes->positions[c->bodyA_index] : NULL;
            if (p) *p = V3(v3add(F3(*p), v3scale(corr, -imA)));
        }
        if (imB > 0.0f) {
            Vec3* p = c->bodyB_type == 0 ? &world->spheres->positions[c->bodyB_index] : c->bodyB_type == 1 ? &world->boxes->positions[c->bodyB_index] : NULL;
            if (p) *p = V3(v3add(F3(*p), v3scale(corr, imB)));
        }
    }
}

}  // extern "C"
