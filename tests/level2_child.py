"""Child process of tests/test_gpu_level2_dropin.py: the UNMODIFIED reference
optimizer (from baseline/_ref, the pip-installed copy that travels with the
repo) with its catalog function table swapped for skypilot_b200's, i.e.
INTEGRATION.md "Level 2": every `sky.catalog.<fn>(..., clouds=...)` call of
the reference's clouds / optimizer lands in skypilot_b200/catalog/
<cloud>_catalog.py and from there on the GPU.

    python tests/level2_child.py <catalog name> <out.json> [scenario ...]
"""
import json
import os
import sys
import tempfile
import time

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _REPO)


def main():
    catalog_name, out_path = sys.argv[1], sys.argv[2]
    wanted = set(sys.argv[3:])
    from tests import scenarios
    from oracle.ref_harness import bootstrap, run_reference
    from skypilot_b200 import synth
    import skypilot_b200 as skyb

    bootstrap.REFERENCE_ROOT = os.path.join(_REPO, 'baseline', '_ref')
    spec = dict(scenarios.CATALOGS[catalog_name])
    enabled = spec.pop('enabled', None) or spec.get('clouds')
    frames = synth.make_catalogs(**spec)
    home = tempfile.mkdtemp(prefix='level2_home_')
    bootstrap.write_catalogs(home, frames)
    sky = bootstrap.import_reference(home, enabled)

    # our side: the same catalog, resident on the GPU
    store = skyb.catalog.load_frames(frames)
    store.set_accelerator_metadata(synth.accelerator_metadata())
    for cloud, frame in synth.images(frames).items():
        store.set_images(cloud, frame)
    skyb.check.set_enabled_clouds(enabled)

    # the one-line change of INTEGRATION.md, as a monkeypatch
    import sky.catalog as ref_catalog
    calls = {'n': 0}
    by_method = {}
    original = ref_catalog._map_clouds_catalog  # pylint: disable=protected-access

    def dispatch(clouds, method_name, *args, **kwargs):
        if clouds is None:
            clouds = list(enabled)
        names = [clouds] if isinstance(clouds, str) else list(clouds)
        if not all(store.has_cloud(n) for n in names):
            # a cloud without a catalog here (Kubernetes, ...): the reference's
            # own module answers, as it would in a partial swap
            return original(clouds, method_name, *args, **kwargs)
        calls['n'] += 1
        t0 = time.perf_counter()
        try:
            return skyb.catalog._map_clouds_catalog(  # pylint: disable=protected-access
                clouds, method_name, *args, **kwargs)
        finally:
            spent = by_method.setdefault(method_name, [0, 0.0])
            spent[0] += 1
            spent[1] += time.perf_counter() - t0

    ref_catalog._map_clouds_catalog = dispatch  # pylint: disable=protected-access

    out = []
    for sc in scenarios.ALL_SUITES[catalog_name]():
        if wanted and sc['name'] not in wanted:
            continue
        try:
            rec = run_reference.run_scenario(sky, sc)
        except Exception as e:  # pylint: disable=broad-except
            import traceback
            rec = {'name': sc['name'],
                   'error': {'type': type(e).__name__, 'message': str(e),
                             'traceback': traceback.format_exc()[-1500:]}}
        if 'candidates' in rec:
            rec['candidates'] = [[[
                c['cloud'], c['instance_type'], c['region'], c['zone'],
                c['value']
            ] for c in cands] for cands in rec['candidates']]
        out.append(rec)
    with open(out_path, 'w', encoding='utf-8') as f:
        json.dump({'records': out, 'catalog_calls': calls['n'],
                   'by_method': by_method}, f)


if __name__ == '__main__':
    main()
