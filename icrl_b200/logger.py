"""Minimal key/value logger with the surface the learner path uses from stable_baselines3/common/logger.py:
`record`, `dump`, `configure` and `Logger.CURRENT.name_to_value` (read by icrl/icrl.py:212 to build `forward/*`)."""
from collections import OrderedDict


class Logger:
    CURRENT = None
    DEFAULT = None

    def __init__(self, folder=None, output_formats=None):
        self.name_to_value = OrderedDict()
        self.name_to_excluded = OrderedDict()
        self.output_formats = output_formats or []

    def record(self, key, value, exclude=None):
        self.name_to_value[key] = value
        self.name_to_excluded[key] = exclude

    def dump(self, step=0):
        for fmt in self.output_formats:
            fmt.write(self.name_to_value, self.name_to_excluded, step)
        self.name_to_value.clear()
        self.name_to_excluded.clear()


Logger.DEFAULT = Logger.CURRENT = Logger()


def configure(folder=None, format_strings=None):
    Logger.CURRENT = Logger(folder=folder)


def record(key, value, exclude=None):
    Logger.CURRENT.record(key, value, exclude)


def dump(step=0):
    Logger.CURRENT.dump(step)


def get_log_dict():
    return Logger.CURRENT.name_to_value


class HumanOutputFormat:
    """Plain key/value table on a text stream (logger.py HumanOutputFormat: write(key_values, key_excluded, step))."""

    def __init__(self, stream):
        self.stream = stream

    def write(self, key_values, key_excluded=None, step=0):
        rows = []
        for k, v in sorted(key_values.items()):
            if key_excluded and key_excluded.get(k) and "stdout" in str(key_excluded[k]):
                continue
            if hasattr(v, "item") and getattr(v, "ndim", 0) == 0:
                v = v.item()
            rows.append((str(k), "%-8.3g" % v if isinstance(v, float) else str(v)))
        if not rows:
            return
        kw, vw = max(len(k) for k, _ in rows), max(len(v) for _, v in rows)
        bar = "-" * (kw + vw + 7)
        self.stream.write("\n".join([bar] + ["| %-*s | %-*s |" % (kw, k, vw, v) for k, v in rows] + [bar]) + "\n")
        self.stream.flush()
