"""GPU tests: `Optimizer.optimize_job_group` (SAME_INFRA placement of a
JobGroup) against plans of the unmodified reference
(tests/golden/jobgroup_*.json, gen_golden.py --job-groups).

Exact comparison where the reference is deterministic: one job, one common
infra, no common infra (independent fallback), errors. With several common
infras and several jobs the reference returns an arbitrary one (see
Optimizer.optimize_job_group); there the test checks that our choice is one
of the reference's common infras, is shared by all jobs, and is the cheapest.
"""
import json
import os

import pytest

import skypilot_b200 as sky
from skypilot_b200.dag import DagExecution
from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu


def _load(catalog):
    path = os.path.join(runner.GOLDEN_DIR, f'jobgroup_{catalog}.json')
    with open(path, encoding='utf-8') as f:
        payload = json.load(f)
    return payload['catalog'], {r['name']: r for r in payload['records']}


def _cases():
    out = []
    for catalog, suite in scenarios.JOB_GROUP_SUITES.items():
        for case in suite():
            out.append(pytest.param(catalog, case,
                                    id=f'{catalog}:{case["name"]}'))
    return out


def _group_cost(tasks):
    return sum(t.best_resources.get_cost(3600) * t.num_nodes for t in tasks)


@pytest.mark.parametrize('catalog,case', _cases())
def test_job_group(catalog, case):
    spec, records = _load(catalog)
    rec = records[case['name']]
    runner.activate_catalog(dict(spec))
    dag, tasks = runner.build_dag(case)
    dag.name = case['name']
    dag.set_execution(DagExecution.PARALLEL)
    target = (sky.OptimizeTarget.COST if case.get('minimize', 'cost') == 'cost'
              else sky.OptimizeTarget.TIME)
    if 'error' in rec:
        with pytest.raises(sky.exceptions.ResourcesUnavailableError) as e:
            sky.Optimizer.optimize_job_group(dag, target, quiet=True)
        assert str(e.value) == rec['error']['message']
        return
    sky.Optimizer.optimize_job_group(dag, target, quiet=True)
    plan = [runner.res_record(t.best_resources) for t in tasks]
    common = rec.get('common_infras', [])
    if target == sky.OptimizeTarget.TIME and len(common) > 1:
        # every region of the fastest cloud ties on time: same cloud only
        assert [p['cloud'] for p in plan] == [p['cloud'] for p in rec['plan']]
        return
    if len(tasks) == 1 or len(common) <= 1:
        assert plan == rec['plan']
        if common:
            overrides = [[[None if r.cloud is None else str(r.cloud).lower(),
                           r.region] for r in list(t.resources)]
                         for t in tasks]
            assert overrides == rec['overrides']
        return
    infras = {(p['cloud'], p['region']) for p in plan}
    assert len(infras) == 1
    assert list(infras.pop()) in common
    # never worse than the reference's pick (cost objective)
    if target == sky.OptimizeTarget.COST:
        ours = _group_cost(tasks)
        dag_r, tasks_r = runner.build_dag(case)
        for t, p in zip(tasks_r, rec['plan']):
            from skypilot_b200.utils import registry
            t.best_resources = sky.Resources(
                cloud=registry.CLOUD_REGISTRY.from_str(p['cloud']),
                instance_type=p['instance_type'], region=p['region'],
                zone=p['zone'], use_spot=p['use_spot'])
        assert ours <= _group_cost(tasks_r) + 1e-9
