"""ctypes binding of the Ballbox C API (reference ballbox.h:30-297).

The same binding drives every library that exports that API:

* the product, ``libballbox_b200.so`` (CUDA path; see ``ballbox_b200.load()``),
* the CPU checkers under ``oracle/`` (tests and bench baselines only).

Struct layouts mirror ``include/ballbox.h`` field for field.  Host arrays are
exposed as numpy views (no copies), like a C user dereferencing
``world->spheres->positions[i]`` (reference README.md:78).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Quat(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class CollisionContact(C.Structure):
    _fields_ = [("bodyA_type", C.c_int), ("bodyA_index", C.c_int),
                ("bodyB_type", C.c_int), ("bodyB_index", C.c_int),
                ("contact_point", Vec3), ("contact_normal", Vec3),
                ("penetration", C.c_float)]


OnCollision = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, C.POINTER(CollisionContact))
SDFFunction = C.CFUNCTYPE(C.c_float, Vec3)


class SphereSystem(C.Structure):
    _fields_ = [("positions", C.POINTER(Vec3)), ("radii", C.POINTER(C.c_float)),
                ("velocities", C.POINTER(Vec3)), ("forces", C.POINTER(Vec3)),
                ("masses", C.POINTER(C.c_float)), ("angular_velocities", C.POINTER(Vec3)),
                ("torques", C.POINTER(Vec3)), ("inertias", C.POINTER(C.c_float)),
                ("is_static", C.POINTER(C.c_bool)), ("is_sleeping", C.POINTER(C.c_bool)),
                ("sleep_timers", C.POINTER(C.c_float)), ("collision_callbacks", C.POINTER(C.c_void_p)),
                ("count", C.c_int), ("capacity", C.c_int)]


class BoxSystem(C.Structure):
    _fields_ = [("positions", C.POINTER(Vec3)), ("sizes", C.POINTER(Vec3)),
                ("rotations", C.POINTER(Quat)), ("velocities", C.POINTER(Vec3)),
                ("forces", C.POINTER(Vec3)), ("masses", C.POINTER(C.c_float)),
                ("angular_velocities", C.POINTER(Vec3)), ("torques", C.POINTER(Vec3)),
                ("inertias", C.POINTER(Vec3)),
                ("is_static", C.POINTER(C.c_bool)), ("is_sleeping", C.POINTER(C.c_bool)),
                ("sleep_timers", C.POINTER(C.c_float)), ("collision_callbacks", C.POINTER(C.c_void_p)),
                ("count", C.c_int), ("capacity", C.c_int)]


class CollisionCollection(C.Structure):
    _fields_ = [("contacts", C.POINTER(CollisionContact)), ("count", C.c_int), ("capacity", C.c_int)]


class ContactConstraint(C.Structure):
    _fields_ = [("bodyA_type", C.c_int), ("bodyA_index", C.c_int),
                ("bodyB_type", C.c_int), ("bodyB_index", C.c_int),
                ("contact_point", Vec3), ("contact_normal", Vec3), ("penetration", C.c_float),
                ("normal_mass", C.c_float), ("tangent_mass", C.c_float * 2), ("tangent", Vec3 * 2),
                ("normal_impulse", C.c_float), ("tangent_impulse", C.c_float * 2),
                ("position_bias", C.c_float), ("restitution", C.c_float), ("friction", C.c_float)]


class ConstraintCollection(C.Structure):
    _fields_ = [("constraints", C.POINTER(ContactConstraint)), ("count", C.c_int), ("capacity", C.c_int)]


class PhysicsSettings(C.Structure):
    _fields_ = [("linear_damping", C.c_float), ("angular_damping", C.c_float),
                ("restitution", C.c_float), ("friction", C.c_float), ("solver_iterations", C.c_int),
                ("baumgarte_bias", C.c_float), ("allowed_penetration", C.c_float),
                ("velocity_threshold", C.c_float),
                ("sleep_linear_threshold", C.c_float), ("sleep_angular_threshold", C.c_float),
                ("sleep_time_required", C.c_float),
                ("manifold_contact_tolerance", C.c_float), ("manifold_penetration_tolerance", C.c_float),
                ("manifold_max_contacts", C.c_int),
                ("collision_epsilon", C.c_float), ("sdf_normal_epsilon", C.c_float),
                ("vector_normalize_epsilon", C.c_float), ("quaternion_epsilon", C.c_float)]


class PhysicsWorld(C.Structure):
    _fields_ = [("spheres", C.POINTER(SphereSystem)), ("boxes", C.POINTER(BoxSystem)),
                ("collisions", C.POINTER(CollisionCollection)),
                ("constraints", C.POINTER(ConstraintCollection)),
                ("gravity", Vec3), ("dt", C.c_float), ("usesSDF", C.c_bool),
                ("SDFCollider", C.c_void_p), ("settings", PhysicsSettings)]


class BallboxStats(C.Structure):
    _fields_ = [("steps", C.c_longlong), ("body_steps", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("spheres", C.c_int), ("boxes", C.c_int), ("candidate_pairs", C.c_int),
                ("contacts", C.c_int), ("solver_contacts", C.c_int), ("touched_bodies", C.c_int),
                ("levels", C.c_int), ("overflow", C.c_int), ("grid_cells", C.c_int),
                ("reserved", C.c_int * 5)]


# numpy dtype of one ContactConstraint record (104 bytes, reference ballbox.h:119-144)
CONSTRAINT_DTYPE = np.dtype([
    ("a_type", "<i4"), ("a_index", "<i4"), ("b_type", "<i4"), ("b_index", "<i4"),
    ("point", "<f4", 3), ("normal", "<f4", 3), ("penetration", "<f4"),
    ("normal_mass", "<f4"), ("tangent_mass", "<f4", 2), ("tangent", "<f4", (2, 3)),
    ("normal_impulse", "<f4"), ("tangent_impulse", "<f4", 2),
    ("position_bias", "<f4"), ("restitution", "<f4"), ("friction", "<f4")])
CONTACT_DTYPE = np.dtype([
    ("a_type", "<i4"), ("a_index", "<i4"), ("b_type", "<i4"), ("b_index", "<i4"),
    ("point", "<f4", 3), ("normal", "<f4", 3), ("penetration", "<f4")])
assert CONSTRAINT_DTYPE.itemsize == 104 and CONTACT_DTYPE.itemsize == 44

SCENE_README, SCENE_BALLS, SCENE_MIXED, SCENE_SDF, SCENE_RLWORLD = 1, 2, 3, 4, 5
SDF_NONE, SDF_PLANE_Y, SDF_WAVY_DET, SDF_WAVY_LIBM, SDF_SPHERE_IN = 0, 1, 2, 3, 4
SYNC_COMPAT, SYNC_RESIDENT, SYNC_COMPAT_BODIES = 0, 1, 2
DL_BODIES, DL_CONSTRAINTS = 1, 2
SOLVER_EXACT, SOLVER_COLOURED, SOLVER_EXACT_FLOW = 0, 1, 2

WP = C.POINTER(PhysicsWorld)


def _view(ptr, count, dtype, shape_tail=()):
    """numpy view over `count` records behind a ctypes pointer (no copy)."""
    if count <= 0 or not ptr:
        return np.zeros((0,) + tuple(shape_tail), dtype=dtype)
    n = int(count)
    for s in shape_tail:
        n *= s
    itemsize = np.dtype(dtype).itemsize
    buf = (C.c_char * (n * itemsize)).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dtype, count=n).reshape((int(count),) + tuple(shape_tail))


class BallboxError(RuntimeError):
    pass


class BallboxLib:
    """One loaded library exporting the Ballbox API."""

    def __init__(self, path: str, product: bool):
        if not os.path.exists(path):
            raise BallboxError(f"library not found: {path}")
        self.path = path
        self.product = product
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        d = self.dll
        d.CreatePhysicsWorld.restype = WP
        d.CreatePhysicsWorld.argtypes = [C.c_int, C.c_int, C.c_int]
        if hasattr(d, "OracleSetWarmStarting"):          # the C restatement's definition of the warm-start extension
            d.OracleSetWarmStarting.argtypes = [WP, C.c_float]; d.OracleSetWarmStarting.restype = None
        if hasattr(d, "BallboxAddBallJoint"):            # joints extension (include/ballbox_ext.h; sequential definition in oracle/)
            d.BallboxAddSpring.argtypes = [WP, C.c_int, C.c_int, C.c_int, C.c_int, Vec3, Vec3, C.c_float, C.c_float, C.c_float]
            d.BallboxAddBallJoint.argtypes = [WP, C.c_int, C.c_int, C.c_int, C.c_int, Vec3]
            d.BallboxAddHingeJoint.argtypes = [WP, C.c_int, C.c_int, C.c_int, C.c_int, Vec3, Vec3, C.c_float]
            d.BallboxAddDistanceJoint.argtypes = [WP, C.c_int, C.c_int, C.c_int, C.c_int, Vec3, Vec3]
            d.BallboxJointCount.argtypes = [WP]; d.BallboxSpringCount.argtypes = [WP]
            d.BallboxClearJoints.argtypes = [WP]; d.BallboxClearJoints.restype = None
        d.DestroyPhysicsWorld.argtypes = [WP]
        d.DestroyPhysicsWorld.restype = None
        d.PhysicsStep.argtypes = [WP, C.c_float]
        d.PhysicsStep.restype = None
        d.AddSphere.argtypes = [WP, Vec3, C.c_float]
        d.AddSphere.restype = C.c_int
        d.AddBox.argtypes = [WP, Vec3, Vec3, Quat]
        d.AddBox.restype = C.c_int
        for name in ("SetSphereMass", "SetBoxMass"):
            getattr(d, name).argtypes = [WP, C.c_int, C.c_float]
            getattr(d, name).restype = None
        for name in ("SetSphereStatic", "SetBoxStatic"):
            getattr(d, name).argtypes = [WP, C.c_int, C.c_bool]
            getattr(d, name).restype = None
        for name in ("AddSphereForce", "AddSphereTorque", "AddBoxForce", "AddBoxTorque"):
            getattr(d, name).argtypes = [WP, C.c_int, Vec3]
            getattr(d, name).restype = None
        for name in ("SetSphereCollisionCallback", "SetBoxCollisionCallback"):
            getattr(d, name).argtypes = [WP, C.c_int, OnCollision]
            getattr(d, name).restype = None
        d.UseSDF.argtypes = [WP, C.c_void_p]
        d.UseSDF.restype = None
        d.QuatFromAxisAngle.argtypes = [Vec3, C.c_float]
        d.QuatFromAxisAngle.restype = Quat
        d.QuatIdentity.restype = Quat
        d.SetDamping.argtypes = [WP, C.c_float, C.c_float]
        d.SetRestitution.argtypes = [WP, C.c_float]
        d.SetCorrectionParameters.argtypes = [WP, C.c_float, C.c_float]
        d.SetSolverIterations.argtypes = [WP, C.c_int]
        d.SetManifoldSettings.argtypes = [WP, C.c_float, C.c_float, C.c_int]
        d.ResetPhysicsSettingsToDefaults.argtypes = [WP]
        d.BbxSceneBuild.argtypes = [WP, C.c_int, C.c_int, C.c_uint]
        d.BbxSceneBuild.restype = C.c_int
        d.BbxSceneCapacity.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        d.BbxSceneCapacity.restype = None
        d.BbxSceneUseSdf.argtypes = [WP, C.c_int]
        d.BbxSceneUseSdf.restype = None
        d.BbxStateHash.argtypes = [WP]
        d.BbxStateHash.restype = C.c_ulonglong
        d.BallboxBackendName.restype = C.c_char_p
        self.stage_names = ("ApplyForces", "IntegrateVelocities", "CollectCollisions", "GenerateContactConstraints",
                            "PreSolveConstraints", "SolveVelocityConstraints", "SolveConstraints", "IntegratePositions",
                            "UpdateSleepStates", "CleanupPhysics")
        self.has_stages = all(hasattr(d, n) for n in self.stage_names)
        if self.has_stages:
            for n in self.stage_names:
                getattr(d, n).argtypes = [WP]
                getattr(d, n).restype = None
        if product:
            d.BallboxSetSyncMode.argtypes = [WP, C.c_int]
            d.BallboxSetSyncMode.restype = None
            d.BallboxUpload.argtypes = [WP]
            d.BallboxDownload.argtypes = [WP, C.c_int]
            d.PhysicsStepN.argtypes = [WP, C.c_float, C.c_int]
            d.BallboxSynchronize.argtypes = [WP]
            d.BallboxSetDevice.argtypes = [WP, C.c_int]
            d.BallboxSetStream.argtypes = [WP, C.c_void_p]
            d.BallboxSetStream.restype = None
            d.BallboxGetLastError.argtypes = [WP]
            d.BallboxGetLastError.restype = C.c_char_p
            d.BallboxClearError.argtypes = [WP]
            d.BallboxClearError.restype = None
            d.UseSDFDevice.argtypes = [WP, C.c_int, C.POINTER(C.c_float)]
            d.UseSDFDevice.restype = None
            d.BallboxSetSDFBakeRegion.argtypes = [WP, Vec3, Vec3, C.c_int]
            d.BallboxSetSDFBakeRegion.restype = None
            d.BallboxBakeSDF.argtypes = [WP]; d.BallboxBakeSDF.restype = C.c_int
            d.BallboxSetSDFBakeGradient.argtypes = [WP, C.c_int]; d.BallboxSetSDFBakeGradient.restype = None
            d.BallboxSetWarmStarting.argtypes = [WP, C.c_float]; d.BallboxSetWarmStarting.restype = None
            d.BallboxSdfInUse.argtypes = [WP]; d.BallboxSdfInUse.restype = C.c_int
            d.BallboxSetSolverSchedule.argtypes = [WP, C.c_int]
            d.BallboxSetSolverSchedule.restype = None
            d.BallboxGetStats.argtypes = [WP, C.POINTER(BallboxStats)]
            d.BallboxDeviceCount.restype = C.c_int
            d.BallboxSetStageTiming.argtypes = [WP, C.c_int]
            d.BallboxSetStageTiming.restype = None
            d.BallboxGetStageTiming.argtypes = [WP, C.POINTER(C.c_float), C.POINTER(C.c_int)]
            d.BatchSetStageTiming.argtypes = [C.c_void_p, C.c_int]
            d.BatchSetStageTiming.restype = None
            d.BatchGetStageTiming.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
            d.CreatePhysicsWorldBatch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
            d.CreatePhysicsWorldBatch.restype = C.c_void_p
            d.DestroyPhysicsWorldBatch.argtypes = [C.c_void_p]
            d.DestroyPhysicsWorldBatch.restype = None
            d.BatchWorldCount.argtypes = [C.c_void_p]
            d.BatchWorld.argtypes = [C.c_void_p, C.c_int]
            d.BatchWorld.restype = WP
            d.BatchSetDevice.argtypes = [C.c_void_p, C.c_int]
            d.BatchSetStream.argtypes = [C.c_void_p, C.c_void_p]
            d.BatchSetStream.restype = None
            d.BatchUpload.argtypes = [C.c_void_p]
            d.BatchDownload.argtypes = [C.c_void_p, C.c_int]
            d.PhysicsStepBatch.argtypes = [C.c_void_p, C.c_float, C.c_int]
            d.BatchSynchronize.argtypes = [C.c_void_p]
            d.BatchGetStats.argtypes = [C.c_void_p, C.POINTER(BallboxStats)]
            d.BatchGetLastError.argtypes = [C.c_void_p]
            d.BatchGetLastError.restype = C.c_char_p
        else:
            for name in ("RefPhysicsStepPruned", "OraclePhysicsStepPruned"):
                if hasattr(d, name):
                    getattr(d, name).argtypes = [WP, C.c_float]
                    getattr(d, name).restype = None

    # ------------------------------------------------------------------ worlds
    def backend_name(self) -> str:
        return self.dll.BallboxBackendName().decode()

    def create_world(self, max_spheres, max_boxes, max_collisions) -> "World":
        p = self.dll.CreatePhysicsWorld(max_spheres, max_boxes, max_collisions)
        if not p:
            raise BallboxError("CreatePhysicsWorld returned NULL")
        return World(self, p, owned=True)

    def scene_capacity(self, kind, n):
        s, b, c = C.c_int(), C.c_int(), C.c_int()
        self.dll.BbxSceneCapacity(kind, n, C.byref(s), C.byref(b), C.byref(c))
        return s.value, b.value, c.value

    def build_scene(self, kind, n, seed, max_collisions=None) -> "World":
        s, b, c = self.scene_capacity(kind, n)
        w = self.create_world(s, b, max_collisions or c)
        if self.dll.BbxSceneBuild(w.ptr, kind, n, seed) < 0:
            raise BallboxError("BbxSceneBuild failed")
        return w


class World:
    """A PhysicsWorld* plus numpy views of its host arrays."""

    def __init__(self, lib: BallboxLib, ptr, owned: bool):
        self.lib, self.ptr, self.owned = lib, ptr, owned
        self._keep = []          # callback thunks must outlive the world

    # lifetime -----------------------------------------------------------------
    def destroy(self):
        if self.ptr and self.owned:
            self.lib.dll.DestroyPhysicsWorld(self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def w(self) -> PhysicsWorld:
        return self.ptr.contents

    # body construction --------------------------------------------------------
    def add_sphere(self, pos, radius):
        return self.lib.dll.AddSphere(self.ptr, Vec3(*pos), radius)

    def add_box(self, pos, half, rot=(0.0, 0.0, 0.0, 1.0)):
        return self.lib.dll.AddBox(self.ptr, Vec3(*pos), Vec3(*half), Quat(*rot))

    def quat_axis_angle(self, axis, angle):
        q = self.lib.dll.QuatFromAxisAngle(Vec3(*axis), angle)
        return (q.x, q.y, q.z, q.w)

    def set_gravity(self, g):
        self.w.gravity = Vec3(*g)

    def set_sphere_callback(self, index, fn):
        thunk = OnCollision(fn)
        self._keep.append(thunk)
        self.lib.dll.SetSphereCollisionCallback(self.ptr, index, thunk)

    def set_box_callback(self, index, fn):
        thunk = OnCollision(fn)
        self._keep.append(thunk)
        self.lib.dll.SetBoxCollisionCallback(self.ptr, index, thunk)

    def set_warm_starting(self, factor):
        d = self.lib.dll
        (d.BallboxSetWarmStarting if self.lib.product else d.OracleSetWarmStarting)(self.ptr, factor)

    def use_sdf(self, sdf_id):
        self.lib.dll.BbxSceneUseSdf(self.ptr, sdf_id)

    # joints extension: a body is (type, index) with type 0 sphere, 1 box, -1 the world -----------------------------
    def add_spring(self, a, b, anchor_a, anchor_b, rest, stiffness, damping=0.0):
        r = self.lib.dll.BallboxAddSpring(self.ptr, a[0], a[1], b[0], b[1], Vec3(*anchor_a), Vec3(*anchor_b), rest, stiffness, damping)
        if r < 0: raise BallboxError("BallboxAddSpring failed")
        return r

    def add_ball_joint(self, a, b, anchor):
        r = self.lib.dll.BallboxAddBallJoint(self.ptr, a[0], a[1], b[0], b[1], Vec3(*anchor))
        if r < 0: raise BallboxError("BallboxAddBallJoint failed")
        return r

    def add_distance_joint(self, a, b, anchor_a, anchor_b):
        r = self.lib.dll.BallboxAddDistanceJoint(self.ptr, a[0], a[1], b[0], b[1], Vec3(*anchor_a), Vec3(*anchor_b))
        if r < 0: raise BallboxError("BallboxAddDistanceJoint failed")
        return r

    def clear_joints(self):
        self.lib.dll.BallboxClearJoints(self.ptr)

    def add_hinge_joint(self, a, b, anchor, axis, half_span=0.25):
        r = self.lib.dll.BallboxAddHingeJoint(self.ptr, a[0], a[1], b[0], b[1], Vec3(*anchor), Vec3(*axis), half_span)
        if r < 0: raise BallboxError("BallboxAddHingeJoint failed")
        return r

    # stepping -----------------------------------------------------------------
    def check(self):
        if self.lib.product:
            e = self.lib.dll.BallboxGetLastError(self.ptr)
            if e:
                raise BallboxError(e.decode())

    def step(self, dt=1.0 / 60.0, pruned=False):
        d = self.lib.dll
        if pruned and not self.lib.product:
            (d.RefPhysicsStepPruned if hasattr(d, "RefPhysicsStepPruned") else d.OraclePhysicsStepPruned)(self.ptr, dt)
        else:
            d.PhysicsStep(self.ptr, dt)
        self.check()

    def stage(self, name):
        """Call one of the reference's public stage functions (ballbox.h:217-229)."""
        getattr(self.lib.dll, name)(self.ptr)
        self.check()

    def step_n(self, dt, n):
        rc = self.lib.dll.PhysicsStepN(self.ptr, dt, n)
        self.check()
        return rc

    def set_sync_mode(self, mode):
        self.lib.dll.BallboxSetSyncMode(self.ptr, mode)

    def upload(self):
        self.lib.dll.BallboxUpload(self.ptr)
        self.check()

    def download(self, what=DL_BODIES | DL_CONSTRAINTS):
        self.lib.dll.BallboxDownload(self.ptr, what)
        self.check()

    def synchronize(self):
        self.lib.dll.BallboxSynchronize(self.ptr)
        self.check()

    def stats(self) -> BallboxStats:
        s = BallboxStats()
        self.lib.dll.BallboxGetStats(self.ptr, C.byref(s))
        return s

    STAGES = ("integrate_v", "broadphase", "narrowphase", "presolve_graph", "solve_kernel", "integrate_p_sleep")

    def set_stage_timing(self, on=True):
        self.lib.dll.BallboxSetStageTiming(self.ptr, 1 if on else 0)

    def stage_timing(self):
        """(dict stage -> summed ms, steps covered) since the last query."""
        out = (C.c_float * 6)()
        n = C.c_int()
        self.lib.dll.BallboxGetStageTiming(self.ptr, out, C.byref(n))
        self.check()
        return {k: float(out[i]) for i, k in enumerate(self.STAGES)}, n.value

    def replay_timing(self):
        """(ms of k_replay_staged summed, steps) for the steps covered by the last stage_timing() call."""
        ms, n = C.c_float(), C.c_int()
        self.lib.dll.BallboxGetReplayTiming.argtypes = [WP, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        self.lib.dll.BallboxGetReplayTiming(self.ptr, C.byref(ms), C.byref(n))
        return float(ms.value), n.value

    def set_stream(self, cuda_stream: int):
        self.lib.dll.BallboxSetStream(self.ptr, C.c_void_p(cuda_stream))

    def state_hash(self) -> int:
        return int(self.lib.dll.BbxStateHash(self.ptr))

    # host array views ---------------------------------------------------------
    @property
    def n_spheres(self):
        return self.w.spheres.contents.count

    @property
    def n_boxes(self):
        return self.w.boxes.contents.count

    def sphere_arrays(self):
        s = self.w.spheres.contents
        n = s.count
        return dict(pos=_view(s.positions, n, "<f4", (3,)), radius=_view(s.radii, n, "<f4"),
                    vel=_view(s.velocities, n, "<f4", (3,)), force=_view(s.forces, n, "<f4", (3,)),
                    mass=_view(s.masses, n, "<f4"), angvel=_view(s.angular_velocities, n, "<f4", (3,)),
                    torque=_view(s.torques, n, "<f4", (3,)), inertia=_view(s.inertias, n, "<f4"),
                    static=_view(s.is_static, n, "u1"), sleeping=_view(s.is_sleeping, n, "u1"),
                    timer=_view(s.sleep_timers, n, "<f4"))

    def box_arrays(self):
        b = self.w.boxes.contents
        n = b.count
        return dict(pos=_view(b.positions, n, "<f4", (3,)), size=_view(b.sizes, n, "<f4", (3,)),
                    rot=_view(b.rotations, n, "<f4", (4,)),
                    vel=_view(b.velocities, n, "<f4", (3,)), force=_view(b.forces, n, "<f4", (3,)),
                    mass=_view(b.masses, n, "<f4"), angvel=_view(b.angular_velocities, n, "<f4", (3,)),
                    torque=_view(b.torques, n, "<f4", (3,)), inertia=_view(b.inertias, n, "<f4", (3,)),
                    static=_view(b.is_static, n, "u1"), sleeping=_view(b.is_sleeping, n, "u1"),
                    timer=_view(b.sleep_timers, n, "<f4"))

    def constraints(self) -> np.ndarray:
        """world->constraints[0..count) as a structured array (view)."""
        cc = self.w.constraints.contents
        return _view(cc.constraints, cc.count, CONSTRAINT_DTYPE)

    def collisions(self) -> np.ndarray:
        cc = self.w.collisions.contents
        return _view(cc.contacts, cc.count, CONTACT_DTYPE)

    def snapshot(self):
        """Deep copy of the mutable state, for comparisons."""
        s, b = self.sphere_arrays(), self.box_arrays()
        out = {"s_" + k: v.copy() for k, v in s.items() if k in ("pos", "vel", "angvel", "sleeping", "timer", "force", "torque")}
        out.update({"b_" + k: v.copy() for k, v in b.items() if k in ("pos", "rot", "vel", "angvel", "sleeping", "timer", "force", "torque")})
        return out
