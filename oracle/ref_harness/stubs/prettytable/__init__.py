"""Minimal stand-in for `prettytable` so the reference imports offline.

Test infrastructure only (oracle/): rows are kept, rendering is a plain join.
"""
FRAME, ALL, NONE, HEADER = 0, 1, 2, 3


class PrettyTable:

    def __init__(self, field_names=None, **kwargs):
        del kwargs
        self.field_names = list(field_names or [])
        self.align = {}
        self.rows = []
        self.max_width = {}
        self.border = True

    def add_row(self, row, **kwargs):
        del kwargs
        self.rows.append(list(row))

    def add_rows(self, rows):
        for row in rows:
            self.add_row(row)

    def get_string(self, **kwargs):
        del kwargs
        lines = [' | '.join(str(c) for c in self.field_names)]
        lines += [' | '.join(str(c) for c in row) for row in self.rows]
        return '\n'.join(lines)

    def __str__(self):
        return self.get_string()
