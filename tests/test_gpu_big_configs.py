"""The two configurations the bench numbers are quoted on, against records
of the UNMODIFIED reference (oracle/ref_harness/gen_big_golden.py):

  cfg4  32-task chain x synthetic 1M-row catalog (BASELINE.json configs[3]),
        plus single tasks / short chains on the same catalog;
  cfg5  10 000 seeded single-task DAGs on the cfg2 catalog (configs[4])
        through Optimizer.optimize_batch, on one and on all visible GPUs.
"""
import hashlib
import json

import pytest

from skypilot_b200 import _native
from skypilot_b200 import engine
from skypilot_b200 import workloads
import skypilot_b200 as sky
from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cfg4():
    payload = runner.load_golden('cfg4_1m')
    assert payload['catalog'] == scenarios.CATALOGS['cfg4_1m']
    store = runner.activate_catalog(payload['catalog'])
    yield store, {r['name']: r for r in payload['records']}
    store.set_scan_mode('auto')


def _check(records, sc):
    got = runner.run_scenario(sc)
    diffs = runner.compare(records[sc['name']], got)
    assert not diffs, '\n'.join(diffs)


def test_cfg4_chain32_matches_reference(cfg4):
    """The workload bench.py times: plan, ordered candidate tables with
    values, objective and totals are the reference's; the scan that ran is
    the class-table scan."""
    store, records = cfg4
    store.set_scan_mode('auto')
    sc = workloads.chain_scenario(32)
    _check(records, sc)
    assert int(engine.LAST_STATS.scan_form) == 4


@pytest.mark.parametrize('mode', ['auto', 'fast-noprune', 'fast-split',
                                  'queue', 'tile'])
def test_cfg4_suite_matches_reference(cfg4, mode):
    store, records = cfg4
    store.set_scan_mode(mode)
    failures = []
    for sc in scenarios.cfg4_scenarios():
        got = runner.run_scenario(sc)
        diffs = runner.compare(records[sc['name']], got)
        if diffs:
            failures.append((sc['name'], diffs[:3]))
    assert not failures, failures


# --------------------------------------------------------------------------
def _cand_digest(cands):
    h = hashlib.md5()
    for c in cands:
        h.update(json.dumps(list(c[:4])).encode())
    return h.hexdigest()


def _cfg5_compare(payload, scs, dags, out):
    messages = payload['messages']
    bad = []
    for sc, dag, res, rec in zip(scs, dags, out, payload['records']):
        assert rec['name'] == sc['name']
        if 'error' in rec:
            if not isinstance(res, sky.exceptions.ResourcesUnavailableError):
                bad.append((sc['name'], 'reference raises', rec['error'][0]))
            continue
        if isinstance(res, Exception):
            bad.append((sc['name'], 'we raise', str(res)[:200]))
            continue
        r = dag.tasks[0].best_resources
        plan = [str(r.cloud).lower(), r.instance_type, r.region, r.zone]
        if plan != rec['plan']:
            bad.append((sc['name'], plan, rec['plan']))
    del messages
    return bad


@pytest.fixture(scope='module')
def cfg5():
    payload = runner.load_golden('cfg5_50k')
    assert payload['catalog'] == scenarios.CATALOGS['cfg5_50k']
    store = runner.activate_catalog(payload['catalog'])
    n = len(payload['records'])
    scs = workloads.cfg5_scenarios(n)
    return store, payload, scs


def test_cfg5_batch_matches_reference_one_gpu(cfg5):
    store, payload, scs = cfg5
    dags = [workloads.build_dag(sc)[0] for sc in scs]
    out = sky.optimize_batch(dags, devices=[0], return_exceptions=True)
    bad = _cfg5_compare(payload, scs, dags, out)
    assert not bad, (len(bad), bad[:5])


def test_cfg5_batch_matches_reference_all_gpus(cfg5):
    store, payload, scs = cfg5
    devices = list(range(_native.device_count()))
    for d in devices:
        store.handle(d)
    dags = [workloads.build_dag(sc)[0] for sc in scs]
    out = sky.optimize_batch(dags, devices=devices, return_exceptions=True)
    bad = _cfg5_compare(payload, scs, dags, out)
    assert not bad, (len(bad), bad[:5])


def test_cfg5_sample_candidate_tables(cfg5):
    """Every 25th DAG one by one: objective, number of candidates and the
    ordered identities of the candidate table."""
    store, payload, scs = cfg5
    bad = []
    for sc, rec in list(zip(scs, payload['records']))[::25]:
        got = runner.run_scenario(sc)
        if 'error' in rec:
            if got.get('error', {}).get('type') != rec['error'][0]:
                bad.append((sc['name'], 'error', got.get('error')))
            continue
        if 'error' in got:
            bad.append((sc['name'], 'we raise', got['error']))
            continue
        if not runner.close(got['objective'], rec['objective']):
            bad.append((sc['name'], got['objective'], rec['objective']))
        cands = got['candidates'][0]
        if len(cands) != rec['n_cand'] or _cand_digest(cands) != rec['cand_md5']:
            bad.append((sc['name'], 'candidates', len(cands), rec['n_cand']))
    assert not bad, (len(bad), bad[:5])
