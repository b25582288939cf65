"""CUDA path vs the unmodified reference (committed golden fixtures).

For every scenario of tests/scenarios.py the fused device path must return
the reference's plan (cloud, instance type, region, zone: identical), the
ordered candidate tables of every task (identical order and identity, values
within 1e-6 relative) and the objective / total cost / total time.
"""
import json
import os

import pytest

from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu


def _cases():
    out = []
    for catalog in scenarios.SUITES:
        # a missing fixture is a failure (load_golden raises), never a skip
        for sc in scenarios.SUITES[catalog]():
            out.append(pytest.param(catalog, sc, id=f'{catalog}:{sc["name"]}'))
    return out


_golden_cache = {}


def _golden(catalog):
    if catalog not in _golden_cache:
        payload = runner.load_golden(catalog)
        _golden_cache[catalog] = (payload['catalog'], {
            r['name']: r for r in payload['records']
        })
    return _golden_cache[catalog]


@pytest.mark.parametrize('catalog,scenario', _cases())
def test_scenario_matches_reference(catalog, scenario):
    spec, records = _golden(catalog)
    assert spec == scenarios.CATALOGS[catalog], (
        'fixture was generated for a different catalog spec; rerun '
        'oracle/ref_harness/gen_golden.py')
    golden = records[scenario['name']]
    runner.activate_catalog(spec)
    got = runner.run_scenario(scenario)
    unordered = any(
        t.get('resources_kind') == 'set' for t in scenario['tasks'])
    diffs = runner.compare(golden, got, unordered_candidates=unordered)
    if diffs:
        out = os.path.join(os.path.dirname(runner.GOLDEN_DIR), '..',
                           'gpurun_out')
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'parity_failures.jsonl'), 'a',
                  encoding='utf-8') as f:
            f.write(json.dumps({'catalog': catalog, 'name': scenario['name'],
                                'diffs': diffs, 'got': got}) + '\n')
    assert not diffs, '\n'.join(diffs)


@pytest.mark.parametrize('mode', ['tile', 'stream', 'stream3', 'queue',
                                  'queue32', 'fast', 'fast-noprune',
                                  'fast-split', 'fast-split-noprune'])
@pytest.mark.parametrize('catalog', ['multi50k', 'aws50k'])
def test_scan_kernel_variants_agree_with_reference(catalog, mode):
    """All scan kernels (one tile per block / TMA streaming with one and
    with three tiles per block / the queue form with few and with 32 tiles
    per block / the class-table scan with and without pruning) must give the
    reference's answers."""
    spec, records = _golden(catalog)
    store = runner.activate_catalog(spec)
    store.set_scan_mode(mode)
    try:
        failures = []
        for sc in scenarios.SUITES[catalog]():
            got = runner.run_scenario(sc)
            unordered = any(
                t.get('resources_kind') == 'set' for t in sc['tasks'])
            diffs = runner.compare(records[sc['name']], got, unordered)
            if diffs:
                failures.append((sc['name'], diffs[:3]))
        assert not failures, failures
    finally:
        store.set_scan_mode('auto')


def test_auto_mode_takes_the_class_table_scan():
    """`auto` must run the fused step_kernel (SkyoptStats.scan_form 4: class-
    table scan, placement and chain DP in one cooperative launch), not
    silently fall back to the round-1 kernels."""
    from skypilot_b200 import engine
    spec, records = _golden('multi6k')
    runner.activate_catalog(spec)
    sc = next(s for s in scenarios.SUITES['multi6k']()
              if s['name'] == 'chain3_mixed')
    got = runner.run_scenario(sc, with_candidates=False)
    assert not runner.compare(records[sc['name']], got)
    assert engine.LAST_STATS is not None
    assert int(engine.LAST_STATS.scan_form) == 4
    assert int(engine.LAST_STATS.total_launches) == 1
