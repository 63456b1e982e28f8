"""cupy.cuda stand-in (see ../__init__.py): only Stream.null.synchronize() is referenced."""


class _NullStream:
    def synchronize(self):
        return None


class Stream:
    null = _NullStream()
