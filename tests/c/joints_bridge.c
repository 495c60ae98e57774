/* A plain C consumer of the joints / springs extension (include/ballbox_ext.h, pulled in by ballbox.h): a plank bridge
 * of boxes tied by hinges between two static posts, a ball hung from the middle plank on a rod, a ball on a spring, and a
 * loose sphere dropped on the bridge.  Prints the state hash and the contact count of every step; tests/test_c_consumer.py
 * builds the same scene through the sequential definition (oracle/bbx_oracle.c) and compares the hashes.
 * Build: gcc -std=gnu11 -Iinclude tests/c/joints_bridge.c -Lballbox_b200 -lballbox_b200 */
#include <stdio.h>
#include "ballbox.h"

unsigned long long BbxStateHash(PhysicsWorld* w);      /* test helper exported next to the scene builders */

#define BODY_SPHERE 0
#define BODY_BOX    1
#define BODY_WORLD  (-1)
#define PLANKS 7

int main(void)
{
    PhysicsWorld* world = CreatePhysicsWorld(16, 16, 512);
    if (!world) { fprintf(stderr, "CreatePhysicsWorld failed\n"); return 2; }
    world->gravity = (Vec3){0, -9.8f, 0};
    int left = AddBox(world, (Vec3){-0.5f, 3, 0}, (Vec3){0.5f, 0.5f, 1}, QuatIdentity());
    int right = AddBox(world, (Vec3){0.5f + 0.8f * PLANKS + 0.5f, 3, 0}, (Vec3){0.5f, 0.5f, 1}, QuatIdentity());
    SetBoxStatic(world, left, true);
    SetBoxStatic(world, right, true);
    int prev = left, mid = -1;
    for (int k = 0; k < PLANKS; k++) {
        int plank = AddBox(world, (Vec3){0.4f + 0.8f * k, 3.3f, 0}, (Vec3){0.38f, 0.06f, 0.8f}, QuatIdentity());
        if (BallboxAddHingeJoint(world, BODY_BOX, prev, BODY_BOX, plank, (Vec3){0.8f * k, 3.3f, 0}, (Vec3){0, 0, 1}, 0.6f) < 0) return 3;
        if (k == PLANKS / 2) mid = plank;
        prev = plank;
    }
    if (BallboxAddHingeJoint(world, BODY_BOX, prev, BODY_BOX, right, (Vec3){0.8f * PLANKS, 3.3f, 0}, (Vec3){0, 0, 1}, 0.6f) < 0) return 3;
    int bob = AddSphere(world, (Vec3){0.4f + 0.8f * (PLANKS / 2), 2.0f, 0}, 0.25f);
    if (BallboxAddDistanceJoint(world, BODY_SPHERE, bob, BODY_BOX, mid, (Vec3){0.4f + 0.8f * (PLANKS / 2), 2.0f, 0},
                                (Vec3){0.4f + 0.8f * (PLANKS / 2), 3.24f, 0}) < 0) return 3;
    int sprung = AddSphere(world, (Vec3){-3, 3, 0}, 0.3f);
    if (BallboxAddSpring(world, BODY_SPHERE, sprung, BODY_WORLD, 0, (Vec3){-3, 3, 0}, (Vec3){-3, 5, 0}, 1.5f, 40.0f, 0.5f) < 0) return 3;
    int loose = AddSphere(world, (Vec3){2.1f, 5.0f, 0.1f}, 0.35f);
    (void)loose;
    if (BallboxJointCount(world) != 2 * (PLANKS + 1) + 1 || BallboxSpringCount(world) != 1) return 4;
    for (int i = 0; i < 240; i++) {
        PhysicsStep(world, 1.0f / 60.0f);
        if (BallboxGetLastError(world)) { fprintf(stderr, "%s\n", BallboxGetLastError(world)); return 5; }
        printf("H %d %016llx %d\n", i, BbxStateHash(world), world->constraints->count);
    }
    DestroyPhysicsWorld(world);
    return 0;
}
