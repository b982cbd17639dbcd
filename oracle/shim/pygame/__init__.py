"""Import-time stand-in for pygame (rendering is out of scope; render_mode=None)."""


class Surface:
    def __init__(self, *a, **k):
        pass


SurfaceType = Surface


class _Event:
    EventType = object

    @staticmethod
    def get():
        return []


event = _Event()
QUIT = KEYDOWN = K_l = K_o = K_m = K_k = K_RIGHT = K_LEFT = K_DOWN = K_UP = SRCALPHA = 0
