"""Device footprint and upload time of the bench catalogs (GPU box).

    python tools/catalog_bytes.py
"""
import ctypes
import sys
import time
sys.path.insert(0, '.')
import skypilot_b200 as sky  # noqa: E402
from skypilot_b200 import _native, synth, workloads  # noqa: E402

for name in ('cfg2', 'cfg4'):
    w = workloads.WORKLOADS[name]
    frames = synth.make_catalogs(**w['catalog'])
    t = time.perf_counter()
    store = sky.catalog.load_frames(frames)
    t1 = time.perf_counter()
    handle = store.handle(0)
    t2 = time.perf_counter()
    dev, row = ctypes.c_int64(0), ctypes.c_int64(0)
    _native.check(_native.load().skyopt_catalog_bytes(
        handle, ctypes.byref(dev), ctypes.byref(row)))
    print(f'{name}: {store.n_real_rows} rows ({store.n_rows} padded), ingest '
          f'{t1 - t:.3f} s, upload + layout build {t2 - t1:.3f} s, device '
          f'{dev.value / 1e6:.1f} MB = {dev.value / store.n_real_rows:.0f} B/row')
