"""CPU, world_size 2, gloo: the N>1 path of the batch API -- shard by proof index, verify each shard independently
(here with the oracle standing in for the device), gather verdict bits -- gives the same verdicts as the unsharded run."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import __graft_entry__ as ge
    import ref_cpu
    from importlib import import_module
    ge.load_package()
    sharding = import_module("snark_bn254_verifier_b200.sharding")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n, rank, world)
    vk, proofs, inputs, expected = ref_cpu.groth16_synth(77, hi - lo, first_index=lo, threads=1)
    _, status = ref_cpu.groth16_verify_batch(vk, proofs, inputs, threads=1)
    assert (status == expected).all()
    verdicts = sharding.gather_verdicts(status, n, dist)
    dist.barrier()
    if rank == 0:
        q.put(verdicts.tolist())
    dist.destroy_process_group()


def test_two_rank_sharded_verdicts_match_unsharded():
    import ref_cpu
    n = 21  # odd: uneven shards, partial last byte of the bitmap
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vk, proofs, inputs, expected = ref_cpu.groth16_synth(77, n, threads=2)
    _, status = ref_cpu.groth16_verify_batch(vk, proofs, inputs, threads=2)
    assert got == (status == 0).tolist() == (expected == 0).tolist()


def test_shard_bounds_cover_and_match_library_rule(pkg):
    from importlib import import_module
    sharding = import_module("snark_bn254_verifier_b200.sharding")
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    st = np.array([0, 1, 0, 0, 6, 8, 0, 16, 0], dtype=np.uint8)
    bits = sharding.pack_verdicts(st)
    assert sharding.unpack_verdicts(bits, 9).tolist() == [True, False, True, True, False, False, True, False, True]
