"""CPU tests of the multi-process (one rank per GPU) path with the gloo
backend, world_size 2: DAG sharding, the max-over-ranks reduction and the
collection of plans on rank 0. The device call itself needs a GPU; here every
rank *states* its shard for the device (the host half of optimize_batch) and
the packed problems are compared with a single-process statement."""
import hashlib
import json
import multiprocessing as mp
import os
import socket

import pytest

from skypilot_b200 import sharding


def test_shard_merge_round_trip():
    items = list(range(11))
    for world in (1, 2, 3, 4, 8):
        parts = [sharding.shard(items, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == items
        assert all(sharding.owner(i, world) == r
                   for r, part in enumerate(parts) for i in part)
        assert sharding.merge(parts, len(items)) == items
    with pytest.raises(ValueError):
        sharding.merge([[0, 2], []], 4)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _make_dags(n):
    import skypilot_b200 as sky
    specs = [dict(accelerators='V100'), dict(cpus='8+'),
             dict(accelerators='T4', use_spot=True), dict(memory='32+'),
             dict(accelerators='A100:8'), dict(cpus='4+', memory='16+')]
    dags = []
    for i in range(n):
        with sky.Dag() as dag:
            sky.Task(f't{i}').set_resources(
                sky.Resources(**specs[i % len(specs)]))
        dags.append(dag)
    return dags


def _state(dags):
    """Packed device problem of a list of single-task DAGs -> digest."""
    import networkx as nx
    import skypilot_b200 as sky
    from skypilot_b200 import engine
    from skypilot_b200.optimizer import Optimizer
    store = sky.catalog.get_store()
    b = engine.ProblemBuilder(store)
    for dag in dags:
        graph = nx.DiGraph()
        graph.add_node(dag.tasks[0])
        Optimizer._state_problem(graph, list(dag.tasks), True, [], True,  # pylint: disable=protected-access
                                 builder=b)
    packed = b.pack()
    h = hashlib.sha256()
    h.update(packed.queries[:packed.n_queries].tobytes())
    h.update(packed.slots[:packed.n_slots].tobytes())
    return {'n_queries': int(packed.n_queries), 'n_slots': int(packed.n_slots),
            'n_dags': len(b.dags), 'digest': h.hexdigest()}


def _worker(rank, world, port, n_dags, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    import skypilot_b200 as sky
    from skypilot_b200 import synth
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sky.catalog.load_frames(
            synth.make_catalogs(seed=2, n_rows=3000,
                                clouds=['aws', 'gcp', 'azure']))
        sky.check.set_enabled_clouds(sky.check.ALL_CATALOG_CLOUDS)
        dags = _make_dags(n_dags)
        mine = sharding.shard(dags, rank, world)
        local = [{'index': i, 'task': d.tasks[0].name}
                 for i, d in zip(range(rank, n_dags, world), mine)]
        stated = _state(mine)
        slowest = sharding.max_over_ranks(1.0 + rank, dist)
        gathered = sharding.gather_on_root(local, n_dags, dist)
        parts = [None] * world
        dist.all_gather_object(parts, stated)
        if rank == 0:
            with open(out_path, 'w', encoding='utf-8') as f:
                json.dump({'slowest': slowest, 'gathered': gathered,
                           'stated': parts, 'single': _state(dags),
                           'per_rank_single': [
                               _state(sharding.shard(dags, r, world))
                               for r in range(world)]}, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo(tmp_path):
    world, n_dags = 2, 13
    port = _free_port()
    out_path = str(tmp_path / 'rank0.json')
    ctx = mp.get_context('spawn')
    procs = [ctx.Process(target=_worker,
                         args=(r, world, port, n_dags, out_path))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    with open(out_path, encoding='utf-8') as f:
        res = json.load(f)
    # the slowest rank sets the reported time
    assert res['slowest'] == 2.0
    # rank 0 sees every DAG's result once, in submission order
    assert [g['index'] for g in res['gathered']] == list(range(n_dags))
    assert [g['task'] for g in res['gathered']] == [
        f't{i}' for i in range(n_dags)]
    # what each rank stated for its GPU is what a single process states for
    # the same shard, and the shards add up to the whole batch
    assert res['stated'] == res['per_rank_single']
    assert sum(s['n_dags'] for s in res['stated']) == n_dags
    assert sum(s['n_queries'] for s in res['stated']) == (
        res['single']['n_queries'])
    assert sum(s['n_slots'] for s in res['stated']) == res['single']['n_slots']
