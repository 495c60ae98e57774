"""Zero-copy batched-worlds interface.

`BatchEnv` owns a PhysicsWorldBatch and exposes its device state as torch tensors that
alias the library's HBM arrays (no copies): observations are read and forces are written
on the device, resets happen on the device, and `step()` never touches the host.
PyTorch is plumbing here (tensor views, streams); every kernel is the library's."""
from __future__ import annotations

import ctypes as C

import torch

from . import binding as bb


class _DeviceViews(C.Structure):
    _fields_ = [("n_worlds", C.c_int), ("cap_spheres", C.c_int), ("cap_boxes", C.c_int),
                ("pos", C.c_void_p), ("vel", C.c_void_p), ("box_rot", C.c_void_p), ("flags", C.c_void_p),
                ("sphere_force", C.c_void_p), ("sphere_torque", C.c_void_p),
                ("box_force", C.c_void_p), ("box_torque", C.c_void_p)]


class _Cai:
    """Minimal __cuda_array_interface__ holder so torch.as_tensor can alias raw device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2, "strides": None}


def _alias(ptr, shape, typestr, device):
    if not ptr or 0 in shape:
        return torch.zeros(shape, dtype=torch.float32 if typestr == "<f4" else torch.uint8, device=device)
    return torch.as_tensor(_Cai(ptr, shape, typestr), device=device)


class BatchEnv:
    def __init__(self, lib: bb.BallboxLib, n_worlds: int, scene_kind=bb.SCENE_RLWORLD, seed_base=7000, device=0):
        self.lib, self.n_worlds = lib, n_worlds
        d = lib.dll
        d.BatchGetDeviceViews.argtypes = [C.c_void_p, C.POINTER(_DeviceViews)]
        d.BatchAddWrenchDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        d.BatchSnapshot.argtypes = [C.c_void_p]
        d.BatchResetWorlds.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        s, b, c = lib.scene_capacity(scene_kind, 0)
        self.batch = d.CreatePhysicsWorldBatch(n_worlds, s, b, c)
        if not self.batch:
            raise bb.BallboxError("CreatePhysicsWorldBatch failed")
        self.worlds = [bb.World(lib, d.BatchWorld(self.batch, i), owned=False) for i in range(n_worlds)]
        for i, w in enumerate(self.worlds):
            d.BbxSceneBuild(w.ptr, scene_kind, 0, seed_base + i)
        self.device = torch.device("cuda", device)
        d.BatchSetDevice(self.batch, device)
        d.BatchSetStream(self.batch, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        self._check(d.BatchUpload(self.batch))
        self._check(d.BatchSnapshot(self.batch))
        v = _DeviceViews()
        self._check(d.BatchGetDeviceViews(self.batch, C.byref(v)))
        W, cs, cb_ = v.n_worlds, v.cap_spheres, v.cap_boxes
        NS, NB = W * cs, W * cb_
        self.cap_spheres, self.cap_boxes = cs, cb_
        pos = _alias(v.pos, (NS + NB, 4), "<f4", self.device)
        vel = _alias(v.vel, (NS + NB, 2, 4), "<f4", self.device)
        self.sphere_pos = pos[:NS].view(W, cs, 4)            # xyz, radius
        self.box_pos = pos[NS:].view(W, cb_, 4)              # xyz, bounding radius
        self.sphere_vel = vel[:NS].view(W, cs, 2, 4)         # [...,0,:3] linear, [...,1,:3] angular
        self.box_vel = vel[NS:].view(W, cb_, 2, 4)
        self.box_rot = _alias(v.box_rot, (NB, 2, 4), "<f4", self.device).view(W, cb_, 2, 4)[:, :, 0, :]
        self.flags = _alias(v.flags, (NS + NB,), "|u1", self.device)

    def _check(self, rc=0):
        e = self.lib.dll.BatchGetLastError(self.batch)
        if e:
            raise bb.BallboxError(e.decode())
        if rc:
            raise bb.BallboxError(f"ballbox call failed with code {rc}")

    def add_wrench(self, sphere_force=None, sphere_torque=None, box_force=None, box_torque=None):
        """Tensors of shape [W, cap, 3] (float32, contiguous, on the device) or None."""
        def ptr(t, cap):
            if t is None:
                return None
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (self.n_worlds, cap, 3)
            return C.c_void_p(t.data_ptr())
        self._check(self.lib.dll.BatchAddWrenchDevice(self.batch, ptr(sphere_force, self.cap_spheres), ptr(sphere_torque, self.cap_spheres),
                                                      ptr(box_force, self.cap_boxes), ptr(box_torque, self.cap_boxes)))

    def step(self, dt=1.0 / 60.0, n=1):
        self._check(self.lib.dll.PhysicsStepBatch(self.batch, dt, n))

    def reset(self, world_ids):
        ids = (C.c_int * len(world_ids))(*world_ids)
        self._check(self.lib.dll.BatchResetWorlds(self.batch, ids, len(world_ids)))

    def download(self, what=bb.DL_BODIES):
        self._check(self.lib.dll.BatchDownload(self.batch, what))

    def close(self):
        if self.batch:
            torch.cuda.synchronize(self.device)
            self.lib.dll.DestroyPhysicsWorldBatch(self.batch)
            self.batch = None
