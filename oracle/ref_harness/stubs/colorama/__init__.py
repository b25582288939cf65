"""Minimal stand-in for `colorama` so the reference imports offline.

Test infrastructure only (oracle/): every colour code is the empty string.
"""


class _Blank:

    def __getattr__(self, name):
        return ''


Fore = _Blank()
Back = _Blank()
Style = _Blank()


def init(*args, **kwargs):
    del args, kwargs
