"""Host-side logic of the z-slab multi-GPU driver, on CPU: the halo-exchange schedule (SlabPlan) is executed with
world_size-2 and world_size-4 gloo process groups on tensors that carry their global plane index, and every rank
checks that it ends up holding exactly the planes z-k / z+k a flood pass with step k reads."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cuda_mesh_voxelization_b200.multi import SlabPlan, cyclic_pieces, parity_boundary, slab_start_buffer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, n, port, plane_elems):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = SlabPlan(n, rank, world)
        T, H = plan.T, plan.H
        # state of plane z is filled with the value z (pass index mixed in so stale halos would be detected)
        for it, k in enumerate(plan.steps()):
            ext = torch.full(((H + T + H) * plane_elems,), -1, dtype=torch.int32)
            center = ext[H * plane_elems:(H + T) * plane_elems]
            center.view(T, plane_elems)[:] = (torch.arange(plan.z0, plan.z1, dtype=torch.int32) + 1000 * it)[:, None]
            far = [torch.full((T * plane_elems,), -1, dtype=torch.int32) for _ in range(2)]
            ops = []
            for t in plan.recvs(k):
                if k < T:
                    lo = (H - k) if t.role == "below" else (H + T)
                    view = ext[lo * plane_elems:(lo + k) * plane_elems]
                else:
                    view = far[0 if t.role == "below" else 1]
                ops.append(dist.P2POp(dist.irecv, view, t.peer))
            for t in plan.sends(k):
                ops.append(dist.P2POp(dist.isend, center[t.src_lo * plane_elems:(t.src_lo + t.count) * plane_elems].clone(), t.peer))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            # what a pass with step k dereferences: plane z-k and z+k for every owned z inside the grid
            for z in range(plan.z0, plan.z1):
                for sgn in (-1, 1):
                    zz = z + sgn * k
                    if zz < 0 or zz >= n:
                        continue
                    if k < T:
                        got = ext.view(H + T + H, plane_elems)[H + (z - plan.z0) + sgn * k]
                    else:
                        got = far[0 if sgn < 0 else 1].view(T, plane_elems)[z - plan.z0]
                    assert torch.all(got == zz + 1000 * it), (rank, k, z, zz, got[:2])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 16), (4, 32), (2, 64)])
def test_halo_exchange_schedule_over_gloo(world, n):
    mp.spawn(_worker, args=(world, n, _free_port(), 8), nprocs=world, join=True)


def test_plan_is_symmetric_and_complete():
    for n, world in [(64, 2), (64, 4), (128, 8), (1024, 8), (1024, 4)]:
        plans = [SlabPlan(n, r, world) for r in range(world)]
        assert [p.z0 for p in plans] == [r * n // world for r in range(world)]
        for k in plans[0].steps():
            sends = {(r, t.peer, t.src_lo, t.count, t.role) for r, p in enumerate(plans) for t in p.sends(k)}
            recvs = {(t.peer, r, t.src_lo, t.count, t.role) for r, p in enumerate(plans) for t in p.recvs(k)}
            assert sends == recvs, (n, world, k)
            for p in plans:  # bytes on the wire per pass: k planes per side below T, whole slabs above
                for t in p.recvs(k):
                    assert t.count == (k if k < p.T else p.T)


def test_bad_partition_is_rejected():
    with pytest.raises(ValueError):
        SlabPlan(100, 0, 3)


def _cyclic_worker(rank, world, n, port, plane_elems):
    """z-cyclic layout -> z-slabs: every rank holds the planes z = rank (mod world), sends the pieces cyclic_pieces() names and
    must end up with exactly the planes of its slab, in order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = n // world
        cyc = torch.arange(rank, n, world, dtype=torch.int32)[:, None].repeat(1, plane_elems)       # plane z holds the value z
        slab = torch.full((T, plane_elems), -1, dtype=torch.int32)
        ops, landing = [], []
        for d, src_first, count, dst_first, dst_stride in cyclic_pieces(rank, world, T):
            piece = cyc[src_first:src_first + count].clone()
            if d == rank:
                slab[dst_first::dst_stride][:count] = piece
            else:
                ops.append(dist.P2POp(dist.isend, piece, d))
        for src in range(world):
            if src == rank:
                continue
            # what rank `src` sends me: its piece for destination `rank`
            (_, _, count, dst_first, dst_stride), = [x for x in cyclic_pieces(src, world, T) if x[0] == rank]
            buf = torch.empty((count, plane_elems), dtype=torch.int32)
            landing.append((buf, dst_first, dst_stride, count))
            ops.append(dist.P2POp(dist.irecv, buf, src))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for buf, dst_first, dst_stride, count in landing:
            slab[dst_first::dst_stride][:count] = buf
        want = torch.arange(rank * T, (rank + 1) * T, dtype=torch.int32)[:, None].repeat(1, plane_elems)
        assert torch.equal(slab, want), (rank, slab[:, 0].tolist())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 16), (4, 64), (8, 128)])
def test_cyclic_to_slab_transpose_over_gloo(world, n):
    mp.spawn(_cyclic_worker, args=(world, n, _free_port(), 4), nprocs=world, join=True)


def test_parity_boundary_planes_cover_the_halos():
    """The two parity launches of a pass together push exactly the nxt lowest and the nxt highest planes of the slab."""
    for T in (16, 128, 512):
        for nxt in (1, 2, 4, 8):
            low, high = set(), set()
            for q in (0, 1):
                (lf, lc), (hf, hc) = parity_boundary(T, nxt, q)
                lo_planes, hi_planes = set(range(lf, lf + 2 * lc, 2)), set(range(hf, hf + 2 * hc, 2))
                assert all(z % 2 == q for z in lo_planes | hi_planes)
                assert not (low & lo_planes) and not (high & hi_planes)
                low |= lo_planes
                high |= hi_planes
            assert low == set(range(nxt)) and high == set(range(T - nxt, T)), (T, nxt, low, high)


def test_slab_phase_start_buffer_leaves_the_sdf_buffer_free():
    """After the z-cyclic phase the slab passes ping-pong between the two extended buffers; whatever the number of GPUs, the
    final pass must read the buffer that does NOT hold the (aliased) signed distance field, i.e. free buffer total % 2."""
    for n in (128, 1024, 2048):
        total = n.bit_length() - 1                       # log2(N) passes
        for world in (2, 4, 8):
            slab_steps = [k for k in (n >> i for i in range(1, total + 1)) if k < world]
            cur = slab_start_buffer(total, len(slab_steps))
            for _ in slab_steps[:-1]:
                cur = 1 - cur                            # a pass reads cur and writes 1 - cur
            assert 1 - cur == total % 2, (n, world)      # the final pass reads cur; the sdf may live in the other buffer

