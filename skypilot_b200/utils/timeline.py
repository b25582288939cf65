"""Timeline of the optimizer path in Chrome's trace-event format -- the hook
of sky/utils/timeline.py:19-140 (`@timeline.event` on `Optimizer.optimize`,
sky/optimizer.py:107, :1036).

Off unless SKYPILOT_TIMELINE_FILE_PATH is set (the reference's variable): then
every decorated call appends a 'B' / 'E' pair, device calls add the kernels'
CUDA-event time as arguments, and the list is written at exit (or by `save`).
With the variable unset a decorated call costs one dictionary look-up.
The native side brackets the same calls with NVTX ranges (`skyopt_optimize`,
`skyopt_scan`, ...), so an Nsight timeline shows host and device together."""
import atexit
import functools
import json
import os
import threading
import time
from typing import Any, Callable, Dict, List, Optional, Union

_events: List[Dict[str, Any]] = []
_lock = threading.Lock()


def file_path() -> Optional[str]:
    return os.environ.get('SKYPILOT_TIMELINE_FILE_PATH')


class Event:
    """`with Event('name'):` or begin() / end()."""

    def __init__(self, name: str, message: Optional[str] = None):
        self._on = bool(file_path())
        self._name = name
        self._message = message

    def _emit(self, phase: str, args: Optional[Dict[str, Any]] = None) -> None:
        event = {
            'name': self._name, 'cat': 'event', 'ph': phase,
            'pid': str(os.getpid()),
            'tid': str(threading.current_thread().ident),
            'ts': f'{time.time() * 10 ** 6: .3f}',
        }
        merged = dict(args or {})
        if self._message is not None:
            merged['message'] = self._message
        if merged:
            event['args'] = merged
        with _lock:
            _events.append(event)

    def begin(self) -> None:
        if self._on:
            self._emit('B')

    def end(self, **args) -> None:
        if self._on:
            self._emit('E', args)

    def __enter__(self) -> 'Event':
        self.begin()
        return self

    def __exit__(self, *exc) -> None:
        self.end()


def event(name_or_fn: Union[str, Callable], message: Optional[str] = None):
    """Decorator: `@event` or `@event('name')`."""

    def wrap(fn: Callable, name: str) -> Callable:

        @functools.wraps(fn)
        def inner(*args, **kwargs):
            if not file_path():
                return fn(*args, **kwargs)
            with Event(name, message):
                return fn(*args, **kwargs)

        return inner

    if callable(name_or_fn):
        fn = name_or_fn
        return wrap(fn, f'{fn.__module__}.{fn.__qualname__}')
    return lambda fn: wrap(fn, name_or_fn)


def device_event(name: str, stats) -> None:
    """A complete ('X') event for one native call from its SkyoptStats."""
    if not file_path() or stats is None:
        return
    total_us = float(stats.total_ms) * 1e3
    now = time.time() * 10**6
    with _lock:
        _events.append({
            'name': name, 'cat': 'device', 'ph': 'X',
            'pid': str(os.getpid()),
            'tid': str(threading.current_thread().ident),
            'ts': f'{now - total_us: .3f}', 'dur': f'{total_us: .3f}',
            'args': {
                'kernels_ms': float(stats.scan_ms) + float(stats.expand_ms) +
                              float(stats.solve_ms),
                'launches': int(stats.total_launches),
                'scan_form': int(stats.scan_form),
                'rows_scanned': int(stats.scan_rows),
            },
        })


def save(path: Optional[str] = None) -> None:
    path = path or file_path()
    if not path:
        return
    with _lock:
        payload = {'traceEvents': list(_events),
                   'displayTimeUnit': 'ms',
                   'otherData': {'source': 'skypilot_b200'}}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'w', encoding='utf-8') as f:
        json.dump(payload, f)


atexit.register(save)
