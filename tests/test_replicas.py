"""CPU test of the N>1 plumbing: world_size-2 gloo processes, max-over-ranks timing and summed throughput."""
import os
import socket
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    rep = ge._sub("replicas")
    r, lr, w = rep.init("gloo")
    assert (r, w) == (rank, world)
    rep.barrier()
    local_ms = 10.0 + 5.0 * rank                 # rank 1 is the slow replica
    rate, t_ms = rep.aggregate_throughput(100.0, local_ms)
    seeds = (rep.replica_seed(7, 0), rep.replica_seed(7, 1))
    out.put((rank, rate, t_ms, rep.max_over_ranks(rank), seeds))
    rep.shutdown()


def test_two_replicas_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rate, t_ms, mx, seeds in res:
        assert t_ms == 15.0                       # max over ranks
        assert abs(rate - 200.0 / 15e-3) < 1e-6   # units of BOTH replicas over the slowest time
        assert mx == 1.0
        assert seeds[0] != seeds[1]


def test_parse_cpulist_and_numa_binding_is_best_effort(pkg):
    import __graft_entry__ as ge
    rep = ge._sub("replicas")
    assert rep.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert rep.parse_cpulist("") == []
    info = rep.bind_to_gpu_numa(0)          # no GPU here: must not raise, must report that nothing was bound
    assert info["bound"] is False
