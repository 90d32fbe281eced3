"""See baseline/stubs/matplotlib/__init__.py."""


def __getattr__(name):
    raise RuntimeError(f"matplotlib.pyplot.{name}: plotting is not available in the benchmark harness (stub module)")
