/*
 * ballbox.h -- drop-in public interface of the B200-native Ballbox PhysicsStep.
 *
 * This header declares the same types (byte-identical layouts) and the same
 * functions as the reference single-header library (reference ballbox.h:30-297)
 * but carries NO implementation: every symbol resolves from libballbox_b200.so,
 * whose PhysicsStep runs as hand-written sm_100a CUDA kernels.  User code that
 * does
 *
 *     #define BALLBOX_IMPLEMENTATION
 *     #include "ballbox.h"
 *
 * keeps compiling unchanged: the macro is accepted and ignored.
 *
 * Layout contract (checked by tests/test_abi.py against the compiled reference):
 *   Vec3 12, Quat 16, CollisionContact 44, ContactConstraint 104,
 *   ContactManifold 160, PhysicsSettings 72, SphereSystem 104, BoxSystem 112,
 *   PhysicsWorld 136 bytes on x86-64.
 *
 * Extensions that the reference does not have (device residency, batches,
 * device SDFs, error and statistics queries) live in ballbox_ext.h.
 */
#ifndef BALLBOX_H
#define BALLBOX_H

#include <stdbool.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHYSICS_PI 3.14159265358979323846f
#define PHYSICS_DEG2RAD (PHYSICS_PI / 180.0f)
#define PHYSICS_RAD2DEG (180.0f / PHYSICS_PI)

#define MAX_MANIFOLD_CONTACTS 8

/* ---- plain data (reference ballbox.h:30-47) ------------------------------ */
typedef struct { float x, y, z; }    Vec3;
typedef struct { float x, y, z, w; } Quat;
typedef struct { Vec3 position; float radius; }            Sphere;
typedef struct { Vec3 position; Vec3 size; Quat rotation; } Box;   /* size = half-extents */

/* body type codes used everywhere below: 0 sphere, 1 box, 2 SDF collider */

/* ---- one contact, normal points from body A to body B (ref :49-57) ------- */
typedef struct {
    int   bodyA_type;
    int   bodyA_index;
    int   bodyB_type;
    int   bodyB_index;
    Vec3  contact_point;
    Vec3  contact_normal;
    float penetration;
} CollisionContact;

/* (ref :60) callbacks are replayed on the host after the device step, in the
 * reference's firing order; the contact pointer is a host copy. */
typedef void (*OnCollision)(int selfIndex, int otherBodyType, int otherIndex,
                            CollisionContact* contact);

/* ---- structure-of-arrays body storage (ref :62-95).  These host arrays are
 * the user-visible mirror of the device state. ------------------------------ */
typedef struct {
    Vec3*        positions;
    float*       radii;
    Vec3*        velocities;
    Vec3*        forces;
    float*       masses;
    Vec3*        angular_velocities;
    Vec3*        torques;
    float*       inertias;            /* scalar, 2/5 m r^2 */
    bool*        is_static;
    bool*        is_sleeping;
    float*       sleep_timers;
    OnCollision* collision_callbacks;
    int          count;
    int          capacity;
} SphereSystem;

typedef struct {
    Vec3*        positions;
    Vec3*        sizes;               /* half-extents */
    Quat*        rotations;
    Vec3*        velocities;
    Vec3*        forces;
    float*       masses;
    Vec3*        angular_velocities;
    Vec3*        torques;
    Vec3*        inertias;            /* diagonal of the body-frame tensor */
    bool*        is_static;
    bool*        is_sleeping;
    float*       sleep_timers;
    OnCollision* collision_callbacks;
    int          count;
    int          capacity;
} BoxSystem;

typedef struct {
    CollisionContact* contacts;
    int count;
    int capacity;
} CollisionCollection;                /* ref :97-101 */

typedef struct {                      /* ref :105-116 */
    Vec3  points[MAX_MANIFOLD_CONTACTS];
    float penetrations[MAX_MANIFOLD_CONTACTS];
    Vec3  normal;
    int   point_count;
    int   bodyA_type;
    int   bodyA_index;
    int   bodyB_type;
    int   bodyB_index;
} ContactManifold;

typedef struct {                      /* ref :119-144 */
    int   bodyA_type;
    int   bodyA_index;
    int   bodyB_type;
    int   bodyB_index;
    Vec3  contact_point;
    Vec3  contact_normal;
    float penetration;
    float normal_mass;
    float tangent_mass[2];
    Vec3  tangent[2];
    float normal_impulse;
    float tangent_impulse[2];
    float position_bias;
    float restitution;
    float friction;
} ContactConstraint;

typedef struct {
    ContactConstraint* constraints;
    int count;
    int capacity;
} ConstraintCollection;               /* ref :146-150 */

typedef float (*SDFFunction)(Vec3 point);   /* ref :153 */

typedef struct {                      /* ref :155-185; defaults: ref :2483-2515 */
    float linear_damping;
    float angular_damping;
    float restitution;
    float friction;
    int   solver_iterations;
    float baumgarte_bias;
    float allowed_penetration;
    float velocity_threshold;
    float sleep_linear_threshold;
    float sleep_angular_threshold;
    float sleep_time_required;
    float manifold_contact_tolerance;
    float manifold_penetration_tolerance;
    int   manifold_max_contacts;
    float collision_epsilon;          /* stored, never read (as in the reference) */
    float sdf_normal_epsilon;
    float vector_normalize_epsilon;   /* stored, never read */
    float quaternion_epsilon;         /* stored, never read */
} PhysicsSettings;

typedef struct {                      /* ref :187-197 */
    SphereSystem*         spheres;
    BoxSystem*            boxes;
    CollisionCollection*  collisions;
    ConstraintCollection* constraints;
    Vec3                  gravity;
    float                 dt;
    bool                  usesSDF;
    SDFFunction           SDFCollider;
    PhysicsSettings       settings;
} PhysicsWorld;

/* ---- world lifetime (ref :200-204) ---------------------------------------- */
PhysicsWorld* CreatePhysicsWorld(int max_spheres, int max_boxes, int max_collisions);
void          DestroyPhysicsWorld(PhysicsWorld* world);
void          UseSDF(PhysicsWorld* world, SDFFunction sdf_function);

/* ---- settings helpers (ref :207-212) -------------------------------------- */
void SetDamping(PhysicsWorld* world, float linear, float angular);
void SetRestitution(PhysicsWorld* world, float restitution);
void SetCorrectionParameters(PhysicsWorld* world, float baumgarte_bias, float allowed_penetration);
void SetSolverIterations(PhysicsWorld* world, int iterations);
void SetManifoldSettings(PhysicsWorld* world, float contact_tolerance,
                         float penetration_tolerance, int max_contacts);
void ResetPhysicsSettingsToDefaults(PhysicsWorld* world);

/* ---- the hot path (ref :215, :1787) ---------------------------------------- */
void PhysicsStep(PhysicsWorld* world, float deltaTime);

/* ---- the stages of PhysicsStep, callable one by one (ref :217-229, order of use :1797-1816).
 * Each call uploads the host arrays, runs that stage on the GPU and downloads what it changed;
 * CollectCollisions leaves the contact list in world->collisions (emission order) and fires the
 * callbacks; the solver stages read world->collisions / world->constraints from the HOST in array
 * order, so lists edited between stages are honoured.  ResolveCollisions and PositionalCorrection
 * are declared but never defined by the reference (:221, :239) and are not provided. */
void ApplyForces(PhysicsWorld* world);
void IntegrateVelocities(PhysicsWorld* world);
void CollectCollisions(PhysicsWorld* world);
void GenerateContactConstraints(PhysicsWorld* world);
void PreSolveConstraints(PhysicsWorld* world);
void SolveVelocityConstraints(PhysicsWorld* world);
void SolveConstraints(PhysicsWorld* world);
void IntegratePositions(PhysicsWorld* world);
void UpdateSleepStates(PhysicsWorld* world);          /* defined (not declared) by the reference, :1696 */
void CleanupPhysics(PhysicsWorld* world);

/* ---- constraint collection management (ref :232-234) ----------------------- */
ConstraintCollection* CreateConstraintCollection(int capacity);
void DestroyConstraintCollection(ConstraintCollection* constraints);
void ClearConstraints(ConstraintCollection* constraints);

/* ---- small math (ref :241-256) --------------------------------------------- */
Vec3  Vec3Add(Vec3 a, Vec3 b);
Vec3  Vec3Sub(Vec3 a, Vec3 b);
Vec3  Vec3Scale(Vec3 v, float s);
float Vec3Dot(Vec3 a, Vec3 b);
Vec3  Vec3Cross(Vec3 a, Vec3 b);
float Vec3Length(Vec3 v);
Vec3  Vec3Normalize(Vec3 v);
Quat  QuatIdentity(void);
Quat  QuatMultiply(Quat q1, Quat q2);
Quat  QuatNormalize(Quat q);
Quat  QuatConjugate(Quat q);
Quat  QuatFromAxisAngle(Vec3 axis, float angle);
Quat  QuatFromEuler(float pitch, float yaw, float roll);
void  QuatToAxisAngle(Quat q, Vec3* axis, float* angle);
Vec3  QuatRotateVec3(Quat q, Vec3 v);

/* ---- spheres (ref :259-268) ------------------------------------------------ */
SphereSystem* CreateSphereSystem(int capacity);
void DestroySphereSystem(SphereSystem* system);
int  AddSphere(PhysicsWorld* world, Vec3 position, float radius);
void UpdateSphere(SphereSystem* system, int index, Vec3 position);            /* defined with external linkage at ref :382 */
void SetSphereMass(PhysicsWorld* world, int index, float mass);
void SetSphereStatic(PhysicsWorld* world, int index, bool is_static);
void AddSphereForce(PhysicsWorld* world, int index, Vec3 force);
void AddSphereTorque(PhysicsWorld* world, int index, Vec3 torque);

/* ---- boxes (ref :271-278) -------------------------------------------------- */
BoxSystem* CreateBoxSystem(int capacity);
void DestroyBoxSystem(BoxSystem* system);
int  AddBox(PhysicsWorld* world, Vec3 position, Vec3 size, Quat rotation);
void UpdateBox(BoxSystem* system, int index, Vec3 position, Quat rotation);   /* defined with external linkage at ref :747 */
void SetBoxMass(PhysicsWorld* world, int index, float mass);
void SetBoxStatic(PhysicsWorld* world, int index, bool is_static);
void AddBoxForce(PhysicsWorld* world, int index, Vec3 force);
void AddBoxTorque(PhysicsWorld* world, int index, Vec3 torque);

/* ---- collision callbacks (ref :281-282) ------------------------------------ */
void SetSphereCollisionCallback(PhysicsWorld* world, int sphereIndex, OnCollision callback);
void SetBoxCollisionCallback(PhysicsWorld* world, int boxIndex, OnCollision callback);

/* ---- standalone queries and helpers outside PhysicsStep (ref :231-239, :262-297).
 * Host functions, as in the reference; their geometry is the __host__ __device__ code the
 * narrowphase kernels run (ballbox_b200/csrc/cuda/bbx_geom.cuh), so a query returns the bits a
 * step produces for that pair.  SDF helpers call the host function given to UseSDF. ------- */
bool CheckSphereCollision(Vec3 pos1, float radius1, Vec3 pos2, float radius2);
bool CheckSphereCollisionWithData(Vec3 pos1, float radius1, Vec3 pos2, float radius2, CollisionContact* contact);
bool CheckSphereSystemCollisions(SphereSystem* system, int sphere_index);
bool CheckBoxSystemCollisions(BoxSystem* system, int box_index);
bool CheckSphereBoxCollision(Vec3 spherePos, float sphereRadius, Vec3 boxPos, Vec3 boxSize, Quat boxRot);
bool CheckSphereBoxCollisionWithData(Vec3 spherePos, float sphereRadius, Vec3 boxPos, Vec3 boxSize, Quat boxRot, CollisionContact* contact);
bool CheckBoxCollision(Vec3 pos1, Vec3 size1, Quat rot1, Vec3 pos2, Vec3 size2, Quat rot2);
bool CheckBoxCollisionWithData(Vec3 pos1, Vec3 size1, Quat rot1, Vec3 pos2, Vec3 size2, Quat rot2, CollisionContact* contact);
void GenerateBoxBoxManifold(PhysicsWorld* world, int boxIndexA, int boxIndexB, ContactManifold* manifold);
void SelectOptimalContactPoints(Vec3* candidate_points, float* penetrations, int candidate_count,
                                ContactManifold* manifold, Vec3 normal, int max_contacts);
void AddManifoldToConstraints(PhysicsWorld* world, ContactManifold* manifold);
Vec3 SDF_EstimateNormal(PhysicsWorld* world, Vec3 point);
bool CheckBoxSDFCollision(PhysicsWorld* world, Vec3 boxPos, Vec3 boxSize, Quat boxRot, CollisionContact* contact);
void GenerateBoxSDFContacts(PhysicsWorld* world, int boxIndex);
/* the pre-constraint impulse path and the position pass: defined by the reference, never called by PhysicsStep */
Vec3  GetBodyVelocityAtPoint(PhysicsWorld* world, int bodyType, int bodyIndex, Vec3 point);
float CalculateEffectiveMass(PhysicsWorld* world, CollisionContact* contact);
void  ApplyImpulse(PhysicsWorld* world, int bodyType, int bodyIndex, Vec3 impulse, Vec3 contactPoint);
void  SolvePositionConstraints(PhysicsWorld* world);

#ifdef __cplusplus
}
#endif

#include "ballbox_ext.h"

#endif /* BALLBOX_H */
