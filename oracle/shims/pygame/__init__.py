"""Minimal headless stand-in for the parts of pygame the RoboRugby reference touches.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  pygame is not installed in the build
container, so the reference package (read-only at /root/reference) is imported against
this shim to generate the golden vectors in tests/golden/.  Only semantics that affect
results are implemented faithfully:

* ``Rect(x, y, w, h)`` truncates every argument toward zero with C ``(int)`` semantics
  (consumed by RR_TrashyPhysics.py:76-85 and RR_EnvBase.py:168-197);
* ``Rect.colliderect`` is the strict-overlap test;
* ``sprite.Group`` is insertion ordered (fixes pair order in RR_TrashyPhysics.py:341-362);
* ``sprite.spritecollide`` lists group members whose ``.rect`` overlaps, self included.

Everything else (display, draw, font, image, transform, time, event, key) is a no-op.
"""
from . import sprite  # noqa: F401

QUIT, KEYDOWN, RLEACCEL, K_ESCAPE = 256, 768, 16384, 27
for _i, _c in enumerate("abcdefghijklmnopqrstuvwxyz"):
    globals()["K_" + _c] = 97 + _i
K_UP, K_DOWN, K_LEFT, K_RIGHT, K_SPACE, K_RETURN = 273, 274, 276, 275, 32, 13


def init():
    return (0, 0)


def quit():
    return None


class Rect:
    __slots__ = ("x", "y", "w", "h")

    def __init__(self, *args):
        if len(args) == 1:
            args = tuple(args[0])
        if len(args) == 2:
            (x, y), (w, h) = args
        else:
            x, y, w, h = args
        self.x, self.y, self.w, self.h = int(x), int(y), int(w), int(h)

    left = property(lambda s: s.x, lambda s, v: setattr(s, "x", int(v)))
    top = property(lambda s: s.y, lambda s, v: setattr(s, "y", int(v)))
    right = property(lambda s: s.x + s.w, lambda s, v: setattr(s, "x", int(v) - s.w))
    bottom = property(lambda s: s.y + s.h, lambda s, v: setattr(s, "y", int(v) - s.h))
    width = property(lambda s: s.w)
    height = property(lambda s: s.h)
    centerx = property(lambda s: s.x + s.w // 2)
    centery = property(lambda s: s.y + s.h // 2)

    @property
    def center(self):
        return (self.centerx, self.centery)

    @center.setter
    def center(self, c):
        self.x = int(c[0]) - self.w // 2
        self.y = int(c[1]) - self.h // 2

    def colliderect(self, o):
        if self.w == 0 or self.h == 0 or o.w == 0 or o.h == 0:
            return False
        return (self.x < o.x + o.w and o.x < self.x + self.w and
                self.y < o.y + o.h and o.y < self.y + self.h)

    def inflate(self, dx, dy):
        return Rect(self.x - dx // 2, self.y - dy // 2, self.w + dx, self.h + dy)

    def copy(self):
        return Rect(self.x, self.y, self.w, self.h)

    def __iter__(self):
        return iter((self.x, self.y, self.w, self.h))

    def __repr__(self):
        return f"<rect({self.x}, {self.y}, {self.w}, {self.h})>"


class Surface:
    def __init__(self, size=(0, 0), *a, **k):
        self._w, self._h = int(size[0]), int(size[1])

    def fill(self, *a, **k):
        return None

    def set_colorkey(self, *a, **k):
        return None

    def convert(self, *a, **k):
        return self

    convert_alpha = convert

    def blit(self, *a, **k):
        return None

    def get_width(self):
        return self._w

    def get_height(self):
        return self._h

    def get_size(self):
        return (self._w, self._h)

    def get_rect(self, **kw):
        r = Rect(0, 0, self._w, self._h)
        for k, v in kw.items():
            setattr(r, k, v)
        return r


class _Noop:
    """Attribute sink: any attribute is a callable returning a harmless value."""

    def __init__(self, ret=None):
        self._ret = ret

    def __getattr__(self, name):
        ret = self._ret
        return lambda *a, **k: ret


class _Display(_Noop):
    def set_mode(self, size=(0, 0), *a, **k):
        return Surface(size)


class _Image(_Noop):
    def load(self, *a, **k):
        return Surface((40, 20))

    def tostring(self, surf, fmt, *a):
        return bytes(surf.get_width() * surf.get_height() * len(fmt))


class _Font(_Noop):
    class _F:
        def render(self, *a, **k):
            return Surface((1, 1))

        def size(self, text):
            return (len(text), 1)

    def SysFont(self, *a, **k):
        return self._F()

    Font = SysFont


class _Transform(_Noop):
    def rotate(self, surf, angle):
        return surf


class _Time(_Noop):
    class Clock:
        def tick(self, *a, **k):
            return 0

    def get_ticks(self):
        return 0


class _Event(_Noop):
    def get(self, *a, **k):
        return []


class _Key(_Noop):
    def get_pressed(self):
        import collections
        return collections.defaultdict(bool)


display = _Display()
image = _Image()
font = _Font()
transform = _Transform()
time = _Time()
event = _Event()
key = _Key()
draw = _Noop()
surface = Surface
