"""Slab domain decomposition across the GPUs of one box: one process per GPU, NCCL over NVLink.

Replaces the reference's shared-memory GridPatch/GhostNode decomposition (Patches/GridPatch.cpp:32-138,
Patches/GhostNode.cpp:58-185; MeshInfo::CreatePatches MeshInfo.cpp:869-1088):

  * 1-D slabs of CELL PLANES along z; rank r holds the particles whose element lies in [cell_lo, cell_hi)
    (the reference's GetPatchForElement, MeshInfo.cpp:1371-1413);
  * after each particle->grid pass the partial sums on the three node planes around an interior slab
    face are swapped with the neighbour and added (the reference's serial ghost->real reductions,
    GhostNode.cpp:127-136,170-185); both sides then own identical complete sums, so the node sweeps run
    redundantly there and no broadcast is needed;
  * after the element reset, particles whose new element left the slab move to the neighbour as packed
    rows (GridPatch::AddMovingParticle / MoveParticlesToNewPatches, GridPatch.cpp:214-251).

libmpmgpu packs/adds the buffers on the device (mpmgpu_slab_* in include/mpmgpu.h); this module moves
them with torch.distributed point-to-point ops.  The exchange logic is written against plain torch
tensors, so the same code runs under gloo with CPU tensors in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

# values per node in a halo exchange: after phases 0, 1, 2 (mass + momentum, force, momentum) and, kind 3, inside
# each XPIC/FMPM iteration (v*next sums)
HALO_VALUES = (4, 3, 3, 3)


def slab_bounds(depth, cell_first, cell_last, world):
    """Split the occupied cell planes [cell_first, cell_last) evenly over `world` ranks; the first and
    last slabs extend to the grid edge.  Returns list of (cell_lo, cell_hi)."""
    n = cell_last - cell_first
    cuts = [cell_first + (n * r) // world for r in range(world + 1)]
    cuts[0] = 0
    cuts[-1] = depth
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def partition_particles(pt, horiz, vert, cell_lo, cell_hi, ids=None):
    """Select the particles whose element z-index lies in [cell_lo, cell_hi); keeps global ids.
    Rigid-BC particles (after n_nonrigid) are few and make node BCs on both sides of a slab face, so every
    rank keeps all of them (they move identically everywhere and never migrate)."""
    in_elem = np.asarray(pt["in_elem"])
    n = in_elem.shape[0]
    n_nr = int(pt.get("n_nonrigid", n))
    k = (in_elem - 1) // (horiz * vert)
    nonrigid = np.arange(n) < n_nr
    sel = np.nonzero(((k >= cell_lo) & (k < cell_hi) & nonrigid) | ~nonrigid)[0]
    out = {}
    for key, v in pt.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[-1] == n:
            out[key] = np.ascontiguousarray(v[..., sel])
        else:
            out[key] = v
    out["ids"] = (np.arange(n, dtype=np.int32) if ids is None else np.asarray(ids, np.int32))[sel]
    out["n_nonrigid"] = int(np.count_nonzero(sel < n_nr))
    return out


class DeviceBuffer:
    """Zero-copy torch view of a raw device pointer owned by libmpmgpu."""

    def __init__(self, ptr, ndoubles):
        self.__cuda_array_interface__ = {"shape": (int(ndoubles),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, ndoubles, device):
    return torch.as_tensor(DeviceBuffer(ptr, ndoubles), device=device)


class NeighbourExchange:
    """Point-to-point exchange with the lower / upper slab neighbour over a torch.distributed group.
    Works on any tensors (CUDA+NCCL in production, CPU+gloo in the tests)."""

    def __init__(self, rank, world, group=None, cpu_group=None):
        self.rank, self.world, self.group = rank, world, group
        self.cpu_group = cpu_group      # optional gloo group for the two-integer handshake (keeps it off the GPU stream)
        self.lower = rank - 1 if rank > 0 else None
        self.upper = rank + 1 if rank < world - 1 else None

    def swap(self, send_lo, send_hi, recv_lo, recv_hi, group=None):
        """Send send_lo to the lower neighbour and send_hi to the upper one; receive theirs."""
        group = group if group is not None else self.group
        ops = []
        if self.lower is not None:
            ops.append(dist.P2POp(dist.isend, send_lo, self.lower, group))
            ops.append(dist.P2POp(dist.irecv, recv_lo, self.lower, group))
        if self.upper is not None:
            ops.append(dist.P2POp(dist.isend, send_hi, self.upper, group))
            ops.append(dist.P2POp(dist.irecv, recv_hi, self.upper, group))
        if not ops:
            return
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def swap_counts(self, n_to_lo, n_to_hi, device):
        """Tell each neighbour how many rows are coming; returns (n_from_lo, n_from_hi).  One small host->device
        copy, one exchange and one device->host read per step (the buffers are allocated once).  With a CPU (gloo)
        group the handshake never touches the GPU stream, so it overlaps whatever kernel is running."""
        if self.cpu_group is not None:
            s_lo = torch.tensor([n_to_lo], dtype=torch.int64)
            s_hi = torch.tensor([n_to_hi], dtype=torch.int64)
            r_lo = torch.zeros(1, dtype=torch.int64)
            r_hi = torch.zeros(1, dtype=torch.int64)
            self.swap(s_lo, s_hi, r_lo, r_hi, group=self.cpu_group)
            return int(r_lo[0]), int(r_hi[0])
        if getattr(self, "_cnt", None) is None or self._cnt[0].device != torch.device(device):
            pin = torch.device(device).type == "cuda"
            self._cnt = (torch.zeros(4, dtype=torch.int64, device=device),
                         torch.zeros(2, dtype=torch.int64).pin_memory() if pin else torch.zeros(2, dtype=torch.int64))
        dev, host = self._cnt
        host[0], host[1] = n_to_lo, n_to_hi
        dev[:2].copy_(host, non_blocking=True)
        dev[2:].zero_()
        self.swap(dev[0:1], dev[1:2], dev[2:3], dev[3:4])
        back = dev[2:].tolist()
        return int(back[0]), int(back[1])

    def swap_rows(self, send_lo, send_hi, recv_lo, recv_hi, n_to_lo, n_to_hi, n_from_lo, n_from_hi, row):
        """Variable-length row exchange (only the non-empty directions are posted; both sides know the counts)."""
        ops = []
        if self.lower is not None:
            if n_to_lo:
                ops.append(dist.P2POp(dist.isend, send_lo[: n_to_lo * row], self.lower, self.group))
            if n_from_lo:
                ops.append(dist.P2POp(dist.irecv, recv_lo[: n_from_lo * row], self.lower, self.group))
        if self.upper is not None:
            if n_to_hi:
                ops.append(dist.P2POp(dist.isend, send_hi[: n_to_hi * row], self.upper, self.group))
            if n_from_hi:
                ops.append(dist.P2POp(dist.irecv, recv_hi[: n_from_hi * row], self.upper, self.group))
        if not ops:
            return
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class SlabSim:
    """One rank of a slab-decomposed run.  `prob` describes the WHOLE grid; `particles` are this rank's
    (see partition_particles) with global ids."""

    def __init__(self, prob, particles, cell_lo, cell_hi, rank, world, device=0, capacity_factor=1.3,
                 migration_capacity=1 << 16, sort_interval=0, group=None, min_capacity=1 << 16):
        from .capi import MpmGpu
        self.rank, self.world = rank, world
        self.device = torch.device("cuda", device)
        n = int(particles["n_nonrigid"])
        # (a slab may start empty and fill by migration: room for at least min_capacity particles)
        self.sim = MpmGpu(prob, device=device, kernel_path=2, max_particles=max(int(n * capacity_factor) + 1024, min_capacity),
                          sort_interval=sort_interval, upload=False)
        self.sim.slab_configure(cell_lo, cell_hi, rank > 0, rank < world - 1, migration_capacity)
        self.sim.upload(particles)
        # kernels and NCCL calls are issued on ONE dedicated stream: ordered without host synchronisation
        # (a non-default stream: handle 0 means "the context's own stream" to mpmgpu_set_stream)
        self.stream = torch.cuda.Stream(device=self.device)
        self.sim.set_stream(self.stream.cuda_stream)
        ptrs, plane_nodes = self.sim.slab_halo_buffers()
        hd = 5 * 3 * plane_nodes
        self.halo = [device_tensor(p, hd, self.device) for p in ptrs]          # send_lo, send_hi, recv_lo, recv_hi
        self.plane_nodes = plane_nodes
        mptrs, self.row, self.mig_cap = self.sim.slab_migration_buffers()
        self.mig = [device_tensor(p, self.row * self.mig_cap, self.device) for p in mptrs]
        cpu_group = None
        # production: the library does the exchanges itself over NCCL (mpmgpu_slab_connect / mpmgpu_slab_step); torch.distributed
        # only carries the communicator ids.  MPMGPU_SLAB_TORCH_EXCHANGE=1 keeps the exchange in this module (torch P2P ops on
        # the library's buffers), which is also what the lock-step cluster and the gloo tests on CPU exercise.
        import os
        self.c_exchange = (world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl"
                           and os.environ.get("MPMGPU_SLAB_TORCH_EXCHANGE", "0") != "1")
        if self.c_exchange:
            obj = [self.sim.nccl_unique_ids() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0, group=group)
            self.sim.slab_connect(rank, world, obj[0])
        elif world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl":
            cpu_group = dist.new_group(backend="gloo")      # collective: every rank builds its SlabSim
        self.ex = NeighbourExchange(rank, world, group, cpu_group)
        if world > 1 and not self.c_exchange:
            self.sim.slab_set_halo_callback(self._halo)      # XPIC/FMPM iterations exchange in the middle of a phase
        self._migrated_out = 0
        self._migrated_in = 0

    @property
    def migrated_out(self):
        return self.sim.slab_migrated()[0] if self.c_exchange else self._migrated_out

    @migrated_out.setter
    def migrated_out(self, v):
        self._migrated_out = v

    @property
    def migrated_in(self):
        return self.sim.slab_migrated()[1] if self.c_exchange else self._migrated_in

    @migrated_in.setter
    def migrated_in(self, v):
        self._migrated_in = v

    def _halo(self, which):
        nd = HALO_VALUES[which] * 3 * self.plane_nodes
        s_lo, s_hi, r_lo, r_hi = self.halo
        self.ex.swap(s_lo[:nd], s_hi[:nd], r_lo[:nd], r_hi[:nd])

    def step(self, nsteps=1):
        if self.c_exchange:
            self.sim.slab_step(nsteps)
            return
        with torch.cuda.stream(self.stream):
            for _ in range(nsteps):
                for phase in range(3):
                    self.sim.slab_phase(phase)
                    if self.world > 1:
                        self._halo(phase)
                # phase 3 = node sweep, element reset (leavers listed), counts on their way to the host, second
                # strain update: the handshake below runs while that last kernel is still busy
                self.sim.slab_phase(3)
                self._migrate()

    def _migrate(self):
        n_lo, n_hi = self.sim.slab_migration_counts()           # waits for the reset kernel only; raises on NaN / overflow
        if self.world == 1:
            return
        f_lo, f_hi = self.ex.swap_counts(n_lo, n_hi, self.device)
        if n_lo or n_hi:
            self.sim.slab_pack_migrants()
        if n_lo or n_hi or f_lo or f_hi:
            s_lo, s_hi, r_lo, r_hi = self.mig
            self.ex.swap_rows(s_lo, s_hi, r_lo, r_hi, n_lo, n_hi, f_lo, f_hi, self.row)
            self.sim.slab_finish_migration(f_lo, f_hi)       # stream-ordered after the exchange: no host sync
            self.migrated_out += n_lo + n_hi
            self.migrated_in += f_lo + f_hi

    def num_particles(self):
        return self.sim.num_particles()

    def download(self):
        return self.sim.download()

    def close(self):
        self.sim.close()


def assemble_by_id(parts, n_global, n_rigid=0):
    """Per-slab downloads (dicts with 'ids') -> global arrays.  The last n_rigid ids are the replicated
    rigid particles: taken from the first slab."""
    out = {}
    for key, v in parts[0].items():
        if key != "ids":
            out[key] = np.zeros(v.shape[:-1] + (n_global,), dtype=v.dtype)
    seen = np.zeros(n_global, dtype=np.int32)
    for r, part in enumerate(parts):
        ids = np.asarray(part["ids"])
        keep = np.ones(ids.shape[0], bool) if r == 0 else ids < n_global - n_rigid
        seen[ids[keep]] += 1
        for key in out:
            out[key][..., ids[keep]] = part[key][..., keep]
    assert np.all(seen == 1), "particle ids lost or duplicated across slabs"
    return out


def gather_by_id(local, n_global, group=None, n_rigid=0):
    """Assemble per-rank downloads into global arrays on every rank (test helper)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    parts = [None] * world
    if world > 1:
        dist.all_gather_object(parts, local, group=group)
    else:
        parts = [local]
    return assemble_by_id(parts, n_global, n_rigid)


class LockstepCluster:
    """All slabs of a run inside ONE process on ONE GPU, stepped in lockstep with device-to-device
    copies standing in for the NCCL exchange.  Same library calls, same buffers, same order as the
    multi-process run: used to debug and to test the slab logic where only one GPU is available."""

    def __init__(self, prob, bounds, device=0, **kw):
        self.sims = []
        world = len(bounds)
        n = prob.nparticles
        for r, (lo, hi) in enumerate(bounds):
            part = partition_particles(prob.particles, prob.horiz, prob.vert, lo, hi)
            self.sims.append(SlabSim(prob, part, lo, hi, r, world, device=device, **kw))
        self.n_global = n
        self.n_rigid = n - int(prob.particles.get("n_nonrigid", n))
        self.world = world

    def _swap_halo(self, which):
        torch.cuda.synchronize()                      # each slab packs on its own stream; the copies below run on torch's
        for r in range(self.world - 1):
            a, b = self.sims[r], self.sims[r + 1]
            nd = HALO_VALUES[which] * 3 * a.plane_nodes
            b.halo[2][:nd].copy_(a.halo[1][:nd])      # a.send_hi -> b.recv_lo
            a.halo[3][:nd].copy_(b.halo[0][:nd])      # b.send_lo -> a.recv_hi
        torch.cuda.synchronize()

    def step(self, nsteps=1):
        for _ in range(nsteps):
            for phase in range(3):
                for s in self.sims:
                    s.sim.slab_phase(phase)
                self._swap_halo(phase)
            for s in self.sims:
                s.sim.slab_phase(3)
            counts = [s.sim.slab_migration_counts() for s in self.sims]
            for s, (n_lo, n_hi) in zip(self.sims, counts):
                if n_lo or n_hi:
                    s.sim.slab_pack_migrants()
            torch.cuda.synchronize()
            for r, s in enumerate(self.sims):
                f_lo = counts[r - 1][1] if r > 0 else 0
                f_hi = counts[r + 1][0] if r < self.world - 1 else 0
                if f_lo:
                    s.mig[2][: f_lo * s.row].copy_(self.sims[r - 1].mig[1][: f_lo * s.row])
                if f_hi:
                    s.mig[3][: f_hi * s.row].copy_(self.sims[r + 1].mig[0][: f_hi * s.row])
            torch.cuda.synchronize()
            for r, s in enumerate(self.sims):
                f_lo = counts[r - 1][1] if r > 0 else 0
                f_hi = counts[r + 1][0] if r < self.world - 1 else 0
                if counts[r][0] or counts[r][1] or f_lo or f_hi:
                    s.sim.slab_finish_migration(f_lo, f_hi)
                    s.migrated_out += counts[r][0] + counts[r][1]
                    s.migrated_in += f_lo + f_hi

    def download(self):
        return assemble_by_id([s.download() for s in self.sims], self.n_global, self.n_rigid)

    def close(self):
        for s in self.sims:
            s.close()
