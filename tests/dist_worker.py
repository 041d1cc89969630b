"""Worker for the multi-process tests (launched by torch.distributed.run).

 mode "shards" (CPU, gloo): every rank asks the C library for its shard; rank 0 checks that the shards tile the grid.
 mode "parity" (GPU, nccl): every rank builds its shard of the grid, runs one iteration; rank 0 checks J / XC / E_xc
                            against the golden fixture and that all ranks hold identical results."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from common import grid_params, load_golden, system_from_golden  # noqa: E402
from dftcxx_b200 import grid as G  # noqa: E402


def shards():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for nshell in (45, 240, 3840, 51484, 7):
        mine = torch.tensor(G.shard_range(nshell, rank, world), dtype=torch.int64)
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        if rank == 0:
            pos = 0
            for first, count in (t.tolist() for t in allr):
                ok &= first == pos and count >= 0
                pos += count
            ok &= pos == nshell
            counts = [t[1].item() for t in allr]
            ok &= max(counts) - min(counts) <= 1
    # the NCCL id is exchanged as an opaque 128-byte blob through the process group
    box = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    got = [None] * world
    dist.all_gather_object(got, box[0])
    if rank == 0:
        ok &= all(g == got[0] and len(g) == 128 for g in got)
        print("SHARDS_OK" if ok else "SHARDS_FAIL")
    dist.destroy_process_group()


def parity(name):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [G.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    g = load_golden(name)
    mg = G.MolecularGrid(system_from_golden(g), device=local, rank=rank, nranks=world)
    mg.set_grid_parameters(*grid_params(g))
    mg.create_grid(box[0])
    J, XC, exc, nel = mg.iteration(g["P"])  # last collective through ncclAllReduce
    peer = False
    want_peer = os.environ.get("DFTGRID_TEST_PEER", "1") == "1"
    can_p2p = all(torch.cuda.can_device_access_peer(a, b) for a in range(world) for b in range(world) if a != b)
    if want_peer:
        peer = mg.connect_peers(dist)
        for _ in range(3):  # several epochs: the exchange buffers alternate and are re-used
            Jp, XCp, excp, nelp = mg.iteration(g["P"])
        # the NCCL ring/tree and the rank-ordered peer sum may round differently; both are within the parity tolerance
        peer_close = bool(np.max(np.abs(Jp - J)) <= 1e-12 and np.max(np.abs(XCp - XC)) <= 1e-12 and excp == exc and nelp == nel)
        J, XC = Jp, XCp
    else:
        peer_close = True
    # fused Fock build on the sharded grid (graph replay from the third call on): F = 2J + XC, E_J = 2 tr(P J)
    for _ in range(3):
        F, ej, excf, nelf = mg.fock(g["P"])
    fock_ok = bool(np.max(np.abs(F - (2.0 * J + XC))) <= 2e-10 and abs(ej - 2.0 * np.trace(g["P"] @ J)) <= 1e-9 * max(1.0, abs(ej))
                   and excf == exc and nelf == nel)
    # a box whose GPUs can reach each other must really have taken the peer-memory path when it was asked for
    peer_expected = want_peer and can_p2p
    # the sharded points are the matching slice of the single-rank grid
    idx = g["idx"]
    mine = (idx >= mg.point_offset) & (idx < mg.point_offset + mg.nloc)
    ok = np.array_equal(mg.get_positions()[idx[mine] - mg.point_offset], g["pts"][mine])
    ok &= bool(np.max(np.abs(mg.get_densities()[idx[mine] - mg.point_offset] - g["rho"][mine])) <= 1e-12 * np.max(g["rho"]))
    res = torch.tensor(np.concatenate([J.ravel(), XC.ravel(), [exc, nel]]), device="cuda")
    allr = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
    flags = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        same = all(torch.equal(a, allr[0]) for a in allr)
        dJ, dXC = np.max(np.abs(J - g["J"])), np.max(np.abs(XC - g["XC"]))
        good = same and peer_close and fock_ok and peer == peer_expected and flags.item() == 1.0 and dJ <= 1e-10 and dXC <= 1e-10 and abs(exc - float(g["exc"])) <= 1e-10 and abs(nel - float(g["nel"])) <= 1e-9
        print("PARITY_%s world=%d dJ=%.2e dXC=%.2e identical_on_all_ranks=%s shard_ok=%s peer_path=%s peer_expected=%s peer_vs_nccl_ok=%s fock_ok=%s" % (
            "OK" if good else "FAIL", world, dJ, dXC, same, flags.item() == 1.0, peer, peer_expected, peer_close, fock_ok))
    mg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1] == "shards":
        shards()
    else:
        parity(sys.argv[2])
