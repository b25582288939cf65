"""Reference records for the two configurations the bench numbers are quoted
on (BASELINE.json configs[3] and configs[4]) -- from the UNMODIFIED reference.

Build container only (needs /root/reference or baseline/_ref):

    python oracle/ref_harness/gen_big_golden.py cfg4_1m
    python oracle/ref_harness/gen_big_golden.py cfg5_50k [--n 10000] [--procs 7]

cfg4_1m   tests/golden/cfg4_1m.json: full records (ordered candidate tables
          with values, plan, objective, totals) of tests.scenarios.cfg4_scenarios
          -- the 32-task chain bench.py times plus single tasks / short chains
          on the same 1M-row catalog.
cfg5_50k  tests/golden/cfg5_50k.json: one COMPACT record per single-task DAG
          of workloads.cfg5_scenarios (10 000 by default): the plan, the
          objective, the number of candidates and an md5 of their ordered
          identities; error type + message (through a string table). The
          scenarios are cut into `procs` slices, one reference process each.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _REPO)

from tests import scenarios  # noqa: E402  pylint: disable=wrong-import-position


def cand_digest(cands) -> str:
    """md5 of the ordered (cloud, instance type, region, zone) identities."""
    h = hashlib.md5()
    for c in cands:
        h.update(json.dumps(list(c[:4])).encode())
    return h.hexdigest()


def compact(rec, messages):
    out = {'name': rec['name']}
    if 'error' in rec:
        msg = rec['error'].get('message', '')
        if msg not in messages:
            messages[msg] = len(messages)
        out['error'] = [rec['error']['type'], messages[msg]]
        return out
    p = rec['plan'][0]
    out['plan'] = [p['cloud'], p['instance_type'], p['region'], p['zone']]
    out['objective'] = rec['objective']
    cands = [[c['cloud'], c['instance_type'], c['region'], c['zone']]
             for c in rec['candidates'][0]]
    out['n_cand'] = len(cands)
    out['cand_md5'] = cand_digest(cands)
    return out


def run_slice(spec, suite, out_path):
    with tempfile.NamedTemporaryFile('w', suffix='.json', delete=False) as f:
        json.dump(suite, f)
        sc_path = f.name
    cmd = [
        sys.executable,
        os.path.join(_HERE, 'run_reference.py'), '--catalog',
        json.dumps(spec), '--scenarios', sc_path, '--out', out_path
    ]
    return subprocess.Popen(cmd, cwd='/tmp', stdout=subprocess.DEVNULL,
                            stderr=subprocess.DEVNULL), sc_path


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('catalog', choices=['cfg4_1m', 'cfg5_50k'])
    parser.add_argument('--n', type=int, default=10000)
    parser.add_argument('--procs', type=int, default=7)
    args = parser.parse_args()
    spec = dict(scenarios.CATALOGS[args.catalog])
    out_dir = os.path.join(_REPO, 'tests', 'golden')
    out_path = os.path.join(out_dir, f'{args.catalog}.json')
    payload = {
        'catalog': spec,
        'generated_by': 'oracle/ref_harness/gen_big_golden.py',
        'reference': 'skypilot-org/skypilot @ 7808630 (unmodified)',
    }
    if args.catalog == 'cfg4_1m':
        suite = scenarios.cfg4_scenarios()
        # one process per scenario slice: the chain alone takes minutes
        slices = [suite[:1], suite[1:]]
    else:
        suite = scenarios.cfg5_scenarios(args.n)
        k = args.procs
        slices = [suite[i::k] for i in range(k)]
        payload['n'] = args.n
    procs = []
    for i, sl in enumerate(slices):
        tmp_out = f'{out_path}.part{i}'
        procs.append((run_slice(spec, sl, tmp_out), tmp_out))
    by_name = {}
    for (proc, sc_path), tmp_out in procs:
        rc = proc.wait()
        os.unlink(sc_path)
        if rc != 0:
            raise SystemExit(f'{args.catalog}: reference harness failed')
        with open(tmp_out, encoding='utf-8') as f:
            for rec in json.load(f):
                by_name[rec['name']] = rec
        os.unlink(tmp_out)
    records = [by_name[sc['name']] for sc in suite]
    if args.catalog == 'cfg4_1m':
        for rec in records:
            if 'candidates' in rec:
                rec['candidates'] = [[[
                    c['cloud'], c['instance_type'], c['region'], c['zone'],
                    c['value']
                ] for c in cands] for cands in rec['candidates']]
        payload['records'] = records
    else:
        messages = {}
        payload['records'] = [compact(r, messages) for r in records]
        payload['messages'] = sorted(messages, key=messages.get)
    with open(out_path, 'w', encoding='utf-8') as f:
        json.dump(payload, f, separators=(',', ':'), sort_keys=True)
    n_err = sum('error' in r for r in payload['records'])
    print(f'[gen_big_golden] wrote {out_path}: {len(records)} records, '
          f'{n_err} errors')


if __name__ == '__main__':
    main()
