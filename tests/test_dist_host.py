"""CPU tests of the row-partitioned path's host logic, including a world_size-2 gloo run
(the N>1 plumbing without GPUs): partitions, halo patterns derived from the slab generator,
and the unique-id broadcast the NCCL communicator uses."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import amg_ann_b200 as ab
from amg_ann_b200 import dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_rows_tiles_the_range():
    for n, k in [(10, 3), (7, 7), (5, 8), (1030301, 8), (100544625, 8)]:
        st = dist.partition_rows(n, k)
        assert st[0] == 0 and st[-1] == n and len(st) == k + 1
        sizes = np.diff(st)
        assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1
    st = dist.slab_partition(464, 8)
    assert st[-1] == 465 ** 3 and all(s % (465 * 465) == 0 for s in st)
    assert (dist.owner_of([0, 4, 7, 10], [0, 3, 4, 9]) == [0, 0, 1, 2]).all()


def test_slab_rows_reference_only_adjacent_slabs_and_cover_the_matrix():
    m, k = 9, 3
    whole = ab.gen.poisson_q1(m, 2, 3, ab.gen.checkerboard_epsv(2, 3, 3.0))
    st = dist.slab_partition(m, k)
    nnz = 0
    for r in range(k):
        sl = ab.gen.poisson_q1(m, 2, 3, ab.gen.checkerboard_epsv(2, 3, 3.0), row_begin=st[r], row_end=st[r + 1])
        b, e = whole.rowptr[st[r]], whole.rowptr[st[r + 1]]
        assert np.array_equal(sl.col, whole.col[b:e]) and np.array_equal(sl.val, whole.val[b:e])
        nnz += sl.nnz
        halo = dist.halo_columns(st, r, sl.col)
        assert set(halo) <= {r - 1, r + 1}
        plane = (m + 1) ** 2
        for q, ids in halo.items():   # one xy-plane on each side (27-point stencil)
            assert len(ids) == plane
            assert ids.min() >= st[q] and ids.max() < st[q + 1]
    assert nnz == whole.nnz


def test_dist_entry_points_refuse_to_run_without_a_device():
    with pytest.raises(ab.AmgbError):
        dist.run_local_group(2, lambda rank, comm: None)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as tdist
    import amg_ann_b200 as ab
    from amg_ann_b200 import dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    m = 8
    st = dist.slab_partition(m, world)
    sl = ab.gen.poisson_q1(m, row_begin=st[rank], row_end=st[rank + 1])
    # what bench.py --workload partitioned does on the host before touching the GPU
    halo = dist.halo_columns(st, rank, sl.col)
    need = torch.zeros(world, dtype=torch.int64)
    for q, ids in halo.items():
        need[q] = len(ids)
    table = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    tdist.all_gather(table, need)
    table = torch.stack(table).numpy()          # table[p][q]: what p needs from q
    assert (table == table.T).all()             # symmetric pattern of the structurally symmetric matrix
    assert table[rank][rank] == 0
    # global sizes by all-reduce, as the level statistics are formed
    tot = torch.tensor([sl.nnz, len(sl.rowptr) - 1], dtype=torch.int64)
    tdist.all_reduce(tot)
    assert tot.tolist() == [(3 * m + 1) ** 3, (m + 1) ** 3]
    # the 128-byte id that rank 0 creates is what every rank ends up with
    ident = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    tdist.broadcast(ident, 0)
    assert ident.tolist() == list(range(128))
    # without a GPU the communicator cannot be built: the product fails loudly
    try:
        ab.Context(0)
        raise SystemExit("expected AMGB_ERR_NO_DEVICE")
    except ab.AmgbError as e:
        assert e.status == -2
    tdist.barrier()
    tdist.destroy_process_group()
    print("ok", rank)
""")


def test_world_size_2_gloo_host_logic(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_reference_arm_under_torchrun_prints_once():
    """The driver launches `bench.py --impl reference --gpus 2` under torchrun like the GPU arm:
    rank 0 alone runs and prints the line, the other rank exits 0 without work."""
    import json
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-m", "8",
                        "--cells", "12"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
