class PrettyTable:  # utils/metrics.py:9 of the reference imports it for a parameter table
    def __init__(self, *a, **kw):
        self.rows, self.field_names = [], list(a[0]) if a else []

    def add_row(self, row):
        self.rows.append(list(row))

    def __str__(self):
        return "\n".join(" | ".join(str(c) for c in r) for r in [self.field_names] + self.rows)
