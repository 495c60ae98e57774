/* A plain C consumer of the drop-in library, following the reference's README quick start
 * (reference README.md:37-87): world of capacity (10, 10, 50), one sphere falling on a static
 * ground box, a sideways push for the first 30 steps, a collision callback on the sphere,
 * 1000 steps of 1/60 s.  Exactly as a user of the reference writes it -- the
 * BALLBOX_IMPLEMENTATION define stays (it is a no-op here), only the include path and the
 * link line change.  Instead of the README's human-readable printf this
 * program prints the bits of every float, so that tests/test_c_consumer.py can compare with
 * tests/golden/readme.npz (generated from the unmodified reference) bit for bit.
 * Build: gcc -std=gnu11 -Iinclude tests/c/readme_quickstart.c -Lballbox_b200 -lballbox_b200 */
#include <stdio.h>
#include <string.h>
#define BALLBOX_IMPLEMENTATION
#include "ballbox.h"

unsigned long long BbxStateHash(PhysicsWorld* w);      /* test helper exported next to the scene builders */

static int g_step;

static unsigned bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

static void OnSphereHitsBox(int selfIndex, int otherBodyType, int otherIndex, CollisionContact* contact)
{
    printf("E %d %d %d %d %08x %08x %08x %08x %08x %08x %08x\n", g_step, selfIndex, otherBodyType, otherIndex,
           bits(contact->penetration), bits(contact->contact_normal.x), bits(contact->contact_normal.y), bits(contact->contact_normal.z),
           bits(contact->contact_point.x), bits(contact->contact_point.y), bits(contact->contact_point.z));
}

int main(void)
{
    PhysicsWorld* world = CreatePhysicsWorld(10, 10, 50);
    if (!world) { fprintf(stderr, "CreatePhysicsWorld failed\n"); return 2; }
    world->gravity = (Vec3){0, -9.8f, 0};
    int sphere = AddSphere(world, (Vec3){0, 5, 0}, 1.0f);
    int ground = AddBox(world, (Vec3){0, -1, 0}, (Vec3){10, 1, 10}, QuatIdentity());
    SetBoxStatic(world, ground, true);
    SetSphereCollisionCallback(world, sphere, OnSphereHitsBox);
    for (int i = 0; i < 1000; i++) {
        g_step = i;
        if (i < 30) AddSphereForce(world, sphere, (Vec3){2.0f, 0, 0});
        PhysicsStep(world, 1.0f / 60.0f);
        printf("H %d %016llx %d\n", i, BbxStateHash(world), world->constraints->count);
        if (i % 10 == 0) {
            Vec3 pos = world->spheres->positions[sphere];
            printf("P %d %08x %08x %08x\n", i, bits(pos.x), bits(pos.y), bits(pos.z));
        }
    }
    DestroyPhysicsWorld(world);
    return 0;
}
