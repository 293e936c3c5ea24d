"""TEST STUB of matplotlib.pyplot: every attribute is a callable that does nothing."""


class _Nothing:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Nothing()
